"""All-reduce of the FC1 gradient alone (tools only): torchrun --nproc-per-node N tools/time_allreduce.py"""
import os, torch, torch.distributed as dist
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl")
for mb in (411, 205, 103, 8):
    x = torch.randn(mb * 1024 * 1024 // 4, device="cuda")
    for _ in range(3): dist.all_reduce(x, op=dist.ReduceOp.AVG)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dist.all_reduce(x, op=dist.ReduceOp.AVG)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"all_reduce {mb} MB fp32: {e0.elapsed_time(e1)/10*1000:.0f} us  busbw {mb*1.048576/ (e0.elapsed_time(e1)/10):.0f} GB/s", flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
