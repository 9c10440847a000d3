#!/usr/bin/env python
"""bench.py — DA hot-path train step (BASELINE.json metric "DA train-step img-pairs/s").

One "step" = one pass of the DA hot path over one batch of synthetic source+target pairs at the
reference's DAF Faster R-CNN R50-DC5 topology on 1024x2048 inputs (C5 = [2,2048,64,128] per pair,
512 RoIs per image):

    ImgAlignmentHead + pixel loss (H1,L1) -> RoIAlign 7x7 (R2) -> shared FCs (F1) ->
    InstanceAlignmentHead + CE (I1,L4) -> consistency loss (L7) -> backward of all of it
    (reversed, lambda-scaled gradient into C5; RoIAlign backward; all weight gradients)
    -> [N>1: one NCCL gradient all-reduce] -> SGD step on every parameter of the path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

`value`  : pairs/s with inputs already resident in HBM (CUDA events, max over ranks).
`e2e`    : same step through the public module API with HOST (pinned) inputs: H2D of the features
           and RoIs and D2H of the loss inside the timed region.
`roofline`: the dominant kernel of the step, timed alone with CUDA events.
`cpu_baseline` / `--impl reference`: the oracle port (reference algorithm restated for CPU,
           oracle/) on the box's host cores, on a bounded sample scaled to the full workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOAD = "daf_org_r50dc5_da_hot_path_1024x2048"
C, H, W, STRIDE = 2048, 64, 128, 16
ROIS_PER_IMG = 512
FC_OUT = 1024
ROI_LAYOUT_DEFAULT = "rhwc"      # memory order of the RoI features in the timed step (see --roi-layout)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="umma_bf16", choices=["umma_bf16", "umma_bf16x6", "umma_bf16x3", "simt_f32"])
    ap.add_argument("--repeats", type=int, default=0, help="how many times the K-step timed region is repeated (0 = enough for ~1.5 s under load)")
    ap.add_argument("--no-chain", action="store_true", help="A/B: instance head layer by layer instead of the chain kernel")
    ap.add_argument("--no-chain-feed", action="store_true", help="A/B: the last shared FC as its own launches, not inside the chain kernel")
    ap.add_argument("--no-fused-tail", action="store_true", help="A/B: separate pixel-head / pixel-loss kernels instead of the fused tail")
    ap.add_argument("--no-f32-line", action="store_true", help="skip the nested line on the fp32-class engine (umma_bf16x6)")
    ap.add_argument("--roi-layout", default=ROI_LAYOUT_DEFAULT, choices=["rchw", "rhwc"],
                    help="memory order of the RoI features between RoIAlign and the first shared FC (rhwc = bin-major, [R,7,7,C])")
    ap.add_argument("--pairs-per-gpu", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="daf", choices=["daf", "maf", "fpn"],
                    help="daf (default, the contract line): DAF-Org hot path; maf: auxiliary line for BASELINE config 3 - the three SRM "
                         "image-level heads of MAFasterRCNN on C3/C4/C5 of 1024x2048 pairs (tensor-bound), one pair per GPU")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0)
    ap.add_argument("--grad-sync", default="peer", choices=["peer", "nccl"],
                    help="N>1: 'peer' = FC1's gradient mean + SGD + operand broadcast in one kernel over NVLink peer memory "
                         "(peer.PeerShardedSGD), NCCL all-reduce for the small tensors; 'nccl' = NCCL all-reduce for everything")
    ap.add_argument("--peer-ctas", type=int, default=0, help="grid cap of the peer kernel (0 = two 128-thread CTAs per SM)")
    ap.add_argument("--no-fused-wgrad-sgd", action="store_true", help="N=1: separate weight-gradient and SGD kernels for FC1")
    ap.add_argument("--peer-transport", default="copy", choices=["copy", "stores"])
    ap.add_argument("--peer-publish", default="deferred", choices=["deferred", "in-step"],
                    help="deferred: the pushes of the refreshed bf16 slices are enqueued after the step and run under the next step's "
                         "RoIAlign forward (the next FC1 forward waits for them on the device)")
    ap.add_argument("--peer-reserve-sms", type=int, default=0, help="SMs the persistent kernels leave free while the peer kernel runs")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def make_host_inputs(pairs, seed, dtype):
    """Pinned host tensors: C5 features (NHWC storage, post-ReLU statistics) and per-image proposals."""
    g = torch.Generator().manual_seed(seed)
    n = 2 * pairs
    c5 = torch.relu(torch.randn(n, H, W, C, generator=g)).to(dtype).pin_memory()
    u = torch.rand(n, ROIS_PER_IMG, 4, generator=g)
    x1 = u[..., 0] * (W * STRIDE - 33)
    y1 = u[..., 1] * (H * STRIDE - 33)
    lo, hi = torch.log(torch.tensor(16.0)), torch.log(torch.tensor(512.0))
    w = torch.exp(lo + u[..., 2] * (hi - lo))
    h = torch.exp(lo + u[..., 3] * (hi - lo))
    boxes = torch.stack([x1, y1, torch.clamp(x1 + w, max=float(W * STRIDE)), torch.clamp(y1 + h, max=float(H * STRIDE))], -1)
    return c5, boxes.contiguous().pin_memory()


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import dist as ddist
    import torch.distributed as dist

    rank, local, world = ddist.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    out, act, step_ms, fused = measure_engine(args, args.engine, rank, local, world, sample_clocks=True)
    if rank == 0:
        out["roofline"], out["kernels"] = kernel_rooflines(dev, act, step_ms, fused_step=fused, roi_layout=args.roi_layout)
    if world == 1 and args.engine == "umma_bf16" and not args.no_f32_line:
        # second line of the same workload on the fp32-class tensor-core engine (exact 3-way bf16 split, 6 product terms on
        # tcgen05; the engine that meets the <= 1e-5 parity bar, tests/test_gpu_parity.py) -- nested, ONE JSON line is printed
        torch.cuda.empty_cache()
        f32, _, _, _ = measure_engine(args, "umma_bf16x6", rank, local, world, sample_clocks=False)
        out["f32_engine"] = {k: f32[k] for k in ("value", "unit", "ms_per_step", "dtype", "e2e", "gpu_launches", "steps", "repeats")}
        out["f32_engine"]["engine"] = "umma_bf16x6 (tcgen05, fp32 operands split hi+mid+lo, fp32 RoIAlign path)"
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_reference(args.cpu_budget_s)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        # captured graphs hold NCCL work; tearing the process group down under them can hang, and the
        # process is ending anyway
        os._exit(0)


def measure_engine(args, engine, rank, local, world, sample_clocks):
    """The timed train step on one engine.  Returns (JSON line without roofline / cpu_baseline, activation dtype,
    ms per step, whether FC1's update is fused into its weight-gradient kernel)."""
    import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import dist as ddist, functional as F_, hotpath, optim, peer, _lib
    import torch.distributed as dist

    dev = torch.device("cuda", local)
    uda.set_engine(engine)
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import da_heads as _dh
    _dh.USE_CHAIN, _dh.USE_FUSED_TAIL = not args.no_chain, not args.no_fused_tail
    _dh.USE_CHAIN_FEED = not args.no_chain_feed
    F_.MANAGED_WGRAD.clear()          # optimizer hooks of a previous measurement are keyed by id(weight)
    act = F_.act_dtype()
    pairs = args.pairs_per_gpu

    torch.manual_seed(0)
    # the bin-major RoI tensor is the bf16 tensor-core engine's layout; the fp32-class engines keep the reference's [R,C,7,7]
    roi_layout = args.roi_layout if engine == "umma_bf16" else "rchw"
    model = hotpath.DAFOrgHotPath(C, STRIDE, FC_OUT, roi_layout=roi_layout).to(dev).train()
    params = ddist.trainable_parameters(model, model.unused_parameters())
    sgd = dict(lr=1e-3, momentum=0.9, weight_decay=5e-4)                      # reference recipe (faster_rcnn_r50_daf_c2f.py:8)
    peer_opt, sync_note = None, "none (1 GPU)"
    if world > 1:
        sync_note = "nccl avg fp32, side stream from the weight-gradient kernel on, persistent kernels on SMs-32 meanwhile"
    fuse = []
    if world == 1 and pairs == 1 and engine == "umma_bf16" and not args.no_fused_wgrad_sgd:
        # one GPU: nothing happens between FC1's weight gradient and its update, so the update rides in the epilogue of the
        # weight-gradient kernel (da_conv_backward_weight_sgd) and the 411 MB gradient is never written or re-read
        fuse = [p for p in params if p.numel() >= (1 << 24)]
        sync_note = "none (1 GPU); FC1's SGD update is applied by the epilogue of its weight-gradient kernel"
    if world > 1 and args.grad_sync == "peer" and pairs == 1 and engine == "umma_bf16":
        big = [p for p in params if p.numel() >= (1 << 24)]
        try:
            peer_opt = peer.PeerShardedSGD(big, max_ctas=args.peer_ctas, reserve_sms=args.peer_reserve_sms, transport=args.peer_transport,
                                            deferred_publish=args.peer_publish == "deferred", **sgd)
            params = [p for p in params if all(p is not q for q in big)]
            how = ("copy engines push gradient slices to their owner and the refreshed bf16 slices to every rank, all-local update kernel"
                   if args.peer_transport == "copy" else "one kernel with SM-issued P2P loads/stores")
            sync_note = (f"FC1 ({sum(p.numel() for p in big) / 1e6:.0f} M params): gradient mean + sharded SGD + bf16 operand broadcast "
                         f"over NVLink peer memory ({how}; da_sgd_step_peer; momentum and fp32 master sharded); "
                         "remaining tensors: nccl avg fp32 + multi-tensor SGD")
        except Exception as e:    # CUDA IPC unavailable on this box: the all-NCCL path is the same math
            peer_opt = None
            F_.MANAGED_WGRAD.clear()
            sync_note += f" (peer path unavailable: {type(e).__name__}: {str(e)[:100]})"
    opt = optim.FusedSGD(params, fuse_wgrad=fuse, **sgd)
    reducer = ddist.OverlappedGradAllReduce(params) if world > 1 else None

    # two input sets (alternated); each is > L2 (C5 alone is 67 MB bf16 per pair, FC1's weight 411 MB)
    host = [make_host_inputs(pairs, 1000 * rank + s, act) for s in range(2)]
    resident = [(c5.to(dev), bx.to(dev)) for c5, bx in host]
    h2d_bytes = host[0][0].numel() * host[0][0].element_size() + host[0][1].numel() * 4

    def step_on(c5_dev, boxes_dev):
        total = None
        for p in range(pairs):
            c5 = c5_dev[2 * p:2 * p + 2].permute(0, 3, 1, 2).requires_grad_(True)    # logical NCHW, channels_last storage
            props = [boxes_dev[2 * p], boxes_dev[2 * p + 1]]
            losses = model.forward_train(c5, props, [0, 1])
            loss, _ = hotpath.parse_losses(losses)
            loss.backward()
            total = loss.detach() if total is None else total + loss.detach()
        if reducer is not None:
            reducer()
        opt.step()
        opt.zero_grad(set_to_none=True)
        if peer_opt is not None:
            peer_opt.join()
        return total

    # The whole step (forward, backward, all-reduce excluded, SGD) is captured once into a CUDA graph and
    # replayed: ~150 kernel launches per step are otherwise CPU-launch bound.  Dropout masks still change
    # every replay through the device-resident seed counter (functional.dropout_counter).
    F_.dropout_counter(dev)
    # ONE graph per input buffer: the two resident input sets double as the static inputs of two captures (shared
    # memory pool), so neither the resident nor the end-to-end loop pays a 67 MB device-to-device staging copy.
    graphs, graph_losses, graph_note = None, None, "eager"
    launches_per_replay = [0]

    def eager_step(c5_dev, boxes_dev):
        F_.bump_dropout_counter(dev)
        return step_on(c5_dev, boxes_dev)

    if not args.no_graph:
        try:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for i in range(3):
                    eager_step(*resident[i % 2])
                    if peer_opt is not None:
                        peer_opt.publish()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graphs, graph_losses, pool = [], [], None
            for slot in range(2):
                gph = torch.cuda.CUDAGraph()
                _lib.reset_launch_count()
                with torch.cuda.graph(gph, pool=pool):
                    graph_losses.append(eager_step(*resident[slot]))
                pool = gph.pool()
                graphs.append(gph)
            launches_per_replay[0] = _lib.launch_count()      # libda_b200 kernels inside one replayed step
            graph_note = "cuda_graph"
        except Exception as e:  # capture is an optimisation; the eager path is the same kernels
            graphs, graph_note = None, f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
            torch.cuda.synchronize()
    graph = graphs

    def run_slot(slot):
        """One step on input set `slot` (resident[slot] holds the inputs)."""
        if graphs is None:
            out = eager_step(*resident[slot])
        else:
            graphs[slot].replay()
            out = graph_losses[slot]
        if peer_opt is not None:
            peer_opt.publish()        # deferred pushes of the refreshed FC1 slices (no-op otherwise); part of the timed step
        return out

    def step_resident(i):
        return run_slot(i % 2)

    # e2e: every step's inputs come from pinned host memory.  The copy of step i+1 is issued on a side
    # stream into the other device buffer while step i computes (double buffering); the loss of every step
    # is read back to the host.
    copy_stream = torch.cuda.Stream(device=dev)
    dev_in = resident      # the H2D copies land in the graphs' static input buffers
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    issued = set()

    def prefetch(i):
        if i in issued:
            return
        issued.add(i)
        slot = i % 2
        c5_h, bx_h = host[i % 2]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[slot])          # the step that last used this buffer has finished
            if not os.environ.get("DA_E2E_NOCOPY"):      # attribution experiment only
                dev_in[slot][0].copy_(c5_h, non_blocking=True)
                dev_in[slot][1].copy_(bx_h, non_blocking=True)
            ev_ready[slot].record(copy_stream)

    loss_ring = [torch.zeros(1).pin_memory() for _ in range(2)]
    ev_loss = [torch.cuda.Event() for _ in range(2)]
    seen = [0.0]

    def step_e2e(i):
        """Step i from HOST inputs; the loss of EVERY step is copied to pinned host memory and read by the host -- one step
        late (while step i runs the host waits for, and reads, the loss of step i-1; the closing synchronize of the timed
        region delivers the last one), so the GPU never idles on the host's read."""
        slot = i % 2
        prefetch(i)
        prefetch(i + 1)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev_ready[slot])
        loss = run_slot(slot)
        ev_free[slot].record(cur)
        loss_ring[slot].copy_(loss.reshape(1), non_blocking=True)
        ev_loss[slot].record(cur)
        if i > 0:
            ev_loss[1 - slot].synchronize()                 # step i-1 has finished: its loss is on the host
            seen[0] = float(loss_ring[1 - slot][0])         # the user reads the loss of every step
        return loss_ring[slot]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, clocks_on, repeats):
        """EXACTLY `steps` steps between (barrier + synchronize) on both sides, CUDA events on the launching stream, max over
        ranks -- repeated `repeats` times back to back (the K-step region of this workload lasts ~40 ms, far too short for
        nvidia-smi to sample clocks and throttle reasons under load); the MEDIAN region is reported, every region is listed."""
        for i in range(warmup):
            fn(i)
        barrier()
        sampler = ClockSampler(local) if clocks_on else None
        if sampler:
            sampler.start()
        regions, launches = [], 0
        for _ in range(repeats):
            barrier()
            _lib.reset_launch_count()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                fn(i)
            e1.record()
            barrier()
            regions.append(ddist.max_over_ranks(e0.elapsed_time(e1), dev))
            launches = _lib.launch_count() if graph is None else launches_per_replay[0] * steps
        clocks = sampler.stop() if sampler else None
        return sorted(regions)[len(regions) // 2], launches, clocks, regions

    # size the number of repeats for ~1.5 s under load (bounded), from a short probe
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(4):
        step_resident(i)
    p1.record()
    barrier()
    probe_ms = ddist.max_over_ranks(p0.elapsed_time(p1), dev) / 4
    repeats = args.repeats if args.repeats > 0 else int(min(50, max(1, round(1500.0 / max(probe_ms * args.steps, 1e-3)))))
    ms_res, launches, clocks, regions = timed(step_resident, args.steps, max(args.warmup, 3), sample_clocks, repeats)
    e2e_calls = [0]

    def step_e2e_seq(_):
        e2e_calls[0] += 1
        return step_e2e(e2e_calls[0] - 1)

    ms_e2e, _, _, _ = timed(step_e2e_seq, args.steps, 2, False, max(1, repeats // 2))
    total_pairs = pairs * world * args.steps
    value = total_pairs / (ms_res / 1e3)
    e2e_value = total_pairs / (ms_e2e / 1e3)

    out = {
        "metric": "da_train_step_img_pairs_per_s", "value": round(value, 3), "unit": "img-pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_res / args.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if engine == "umma_bf16" else "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": pairs, "c5": [2 * pairs, C, H, W], "rois_per_img": ROIS_PER_IMG,
                   "engine": engine, "roi_layout": roi_layout, "launch": graph_note, "grad_allreduce": sync_note, "step": "H1+L1, RoIAlign fwd/bwd, shared FCs, I1+L4, L7, backward, SGD",
                   "l2_policy": "inputs_and_weights_exceed_L2 (C5 67MB/pair bf16, FC1 weight 411MB, RoI features 205MB); two input sets alternated"},
        "e2e": {"value": round(e2e_value, 3), "unit": "img-pairs/s", "h2d_bytes_per_step": int(h2d_bytes), "d2h_bytes_per_step": 4,
                "ms_per_step": round(ms_e2e / args.steps, 4),
                "h2d_gbs_per_rank": round(h2d_bytes / (ms_e2e / args.steps) / 1e6, 2)},
        "gpu_launches": int(launches), "clocks": clocks,
        "repeats": repeats, "region_ms": [round(r, 3) for r in regions],
    }
    if peer_opt is not None:
        peer_opt.check_errors()
    return out, act, ms_res / args.steps, bool(fuse)


def kernel_rooflines(dev, act, step_ms, fused_step=True, roi_layout=ROI_LAYOUT_DEFAULT):
    """Per-kernel timings of the step's hot kernels AT THE STEP'S SIZES (one pair: 2 images, 1024 RoIs), each
    alone, CUDA events on the launching stream, L2 flushed between launches.  Algorithmic bytes / flops per
    launch are the DESIGN.md §4 figures.  Returns (roofline of the dominant kernel, table)."""
    import ctypes
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_
    from unsupervised_domain_adaptation_object_detection_implementation_b200._lib import lib, check
    pk = peaks()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(0)
    N, R = 2, 2 * ROIS_PER_IMG
    feat = torch.relu(torch.randn(N, H, W, C, device=dev, generator=g)).to(act).permute(0, 3, 1, 2)
    cpu_g = torch.Generator().manual_seed(1)
    u = torch.rand(R, 4, generator=cpu_g)
    x1, y1 = u[:, 0] * (W * STRIDE - 33), u[:, 1] * (H * STRIDE - 33)
    wh = torch.exp(torch.log(torch.tensor(16.0)) + u[:, 2:] * (torch.log(torch.tensor(512.0)) - torch.log(torch.tensor(16.0))))
    rois = torch.stack([(torch.arange(R) // (R // N)).float(), x1, y1, torch.clamp(x1 + wh[:, 0], max=W * STRIDE),
                        torch.clamp(y1 + wh[:, 1], max=H * STRIDE)], 1).to(dev)
    es = feat.element_size()

    def time_it(fn, iters=10):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return sum(ts[:max(1, len(ts) // 2)]) / max(1, len(ts) // 2)

    table = {}

    def hbm_row(name, t, nbytes, note=None):
        table[name] = {"ms": round(t, 4), "bytes": nbytes, "gbs": round(nbytes / t / 1e6, 1),
                       "frac": round(nbytes / t / 1e6 / pk["hbm_gbs"], 4)}
        if note:
            table[name]["note"] = note

    def tensor_row(name, t, fl, note=None):
        table[name] = {"ms": round(t, 4), "flops": fl, "tflops": round(fl / t / 1e9, 1), "frac": round(fl / t / 1e9 / pk["bf16_tflops"], 4)}
        if note:
            table[name]["note"] = note

    # RoIAlign 7x7 forward / backward: pooled tensor + feature map, each touched once
    for lay in (roi_layout, "rhwc" if roi_layout == "rchw" else "rchw"):
        in_step = lay == roi_layout
        sfx = "" if in_step else "_" + lay
        note = "incl. the RoI prep kernel" + ("" if in_step else f"; RoI tensor in {lay} order (not the step's layout)")
        out = F_.roi_align(feat, rois, 7, 1.0 / STRIDE, out_layout=lay)
        hbm_row("roi_align_fwd" + sfx, time_it(lambda: F_.roi_align(feat, rois, 7, 1.0 / STRIDE, out_layout=lay)),
                R * C * 49 * es + N * C * H * W * es + R * 20, note)
        gout = torch.randn(out.shape, device=dev, generator=g).to(out.dtype)
        fr = feat.detach().requires_grad_(True)
        o2 = F_.roi_align(fr, rois, 7, 1.0 / STRIDE, out_layout=lay)
        hbm_row("roi_align_bwd" + sfx, time_it(lambda: torch.autograd.grad(o2, fr, gout, retain_graph=True)),
                R * C * 49 * es + N * C * H * W * es + R * 20, note)
        del out, o2, gout, fr
    del feat
    # FC1 of the shared head: [1024, 100352] x [100352 -> 1024], forward / data gradient / weight gradient
    K = C * 49
    x = torch.randn(R, 1, 1, K, device=dev, generator=g).to(torch.bfloat16)
    w = (torch.randn(FC_OUT, K, device=dev, generator=g) * K ** -0.5).to(torch.bfloat16)
    dz = torch.randn(R, 1, 1, FC_OUT, device=dev, generator=g).to(torch.bfloat16)
    y = torch.empty(R, 1, 1, FC_OUT, device=dev, dtype=torch.bfloat16)
    dx = torch.empty_like(x)
    dw = torch.empty(FC_OUT, 1, 1, K, device=dev, dtype=torch.float32)
    desc = F_._conv_desc(R, 1, 1, K, FC_OUT, 1, 1, 1, 0, "umma_bf16", torch.bfloat16, torch.bfloat16)
    ws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(desc)), torch.device(dev), "conv")
    P, S = F_._ptr, F_._stream
    fl = 2.0 * R * K * FC_OUT
    tensor_row("fc1_fwd", time_it(lambda: check(lib.da_conv_forward(ctypes.byref(desc), P(x), P(w), None, None, 1, 0.0, 0, P(y),
                                                                     P(ws), ws.numel(), S()), "conv_forward")), fl)
    tensor_row("fc1_dgrad", time_it(lambda: check(lib.da_conv_backward_data(ctypes.byref(desc), P(dz), P(w), 1.0, P(dx), P(ws),
                                                                             ws.numel(), S()), "conv_backward_data")), fl)
    tensor_row("fc1_wgrad", time_it(lambda: check(lib.da_conv_backward_weight(ctypes.byref(desc), P(x), P(dz), P(dw), P(ws),
                                                                               ws.numel(), S()), "conv_backward_weight")), fl,
               "incl. the split-K reduction")
    # the same weight gradient with the SGD step of that weight applied by its epilogue (the N=1 train step uses this one):
    # HBM-bound - master + momentum read, master + momentum + bf16 copy written (18 B per parameter), operands read once
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import _lib as _l
    n = FC_OUT * K
    wf = torch.randn(n, device=dev, generator=g)
    buf = torch.zeros(n, device=dev)
    shadow = w.view(-1)
    rec = _l.SgdFuse(wf.data_ptr(), buf.data_ptr(), shadow.data_ptr(), 0.01, 0.9, 1e-4, 0)
    hbm_row("fc1_wgrad_sgd", time_it(lambda: check(lib.da_conv_backward_weight_sgd(ctypes.byref(desc), P(x), P(dz), ctypes.byref(rec), P(ws),
                                                                                   ws.numel(), S()), "conv_backward_weight_sgd")),
            n * 18 + R * K * 2 + R * FC_OUT * 2, "tcgen05 weight gradient + fused SGD epilogue; 2*M*N*K = 210 GFLOP ride along")
    del x, dx, y, dz
    # ---- metric (iii): the DA convs themselves (tensor-pipe utilisation = achieved / measured bf16 burst peak), at the
    # 1024x2048 sizes: H1 1x1 2048->512 on C5 (resnet_da_daf_org.py:124), SRM 3x3 pad-3 512->4608 on C5's 66x130 conv1
    # output (resnet_da.py:89-91, Q12) and the Global head's 3x3 stride-2 2048->1024 (resnet_da_cbam.py:123)
    def conv_rows(tag, n_, h_, w_, cin, cout, k, stride, pad):
        oh, ow = (h_ + 2 * pad - k) // stride + 1, (w_ + 2 * pad - k) // stride + 1
        cx = (torch.randn(n_, h_, w_, cin, device=dev, generator=g)).to(torch.bfloat16)
        cw = (torch.randn(cout, k, k, cin, device=dev, generator=g) * (cin * k * k) ** -0.5).to(torch.bfloat16)
        cy = torch.empty(n_, oh, ow, cout, device=dev, dtype=torch.bfloat16)
        cdz = torch.randn(n_, oh, ow, cout, device=dev, generator=g).to(torch.bfloat16)
        cdx = torch.empty_like(cx)
        cdw = torch.empty(cout, k, k, cin, device=dev, dtype=torch.float32)
        sc = torch.ones(cout, device=dev)
        cd = F_._conv_desc(n_, h_, w_, cin, cout, k, k, stride, pad, "umma_bf16", torch.bfloat16, torch.bfloat16)
        cws = F_.workspace(lib.da_conv_workspace_bytes(ctypes.byref(cd)), torch.device(dev), "conv")
        cfl = 2.0 * n_ * oh * ow * cout * cin * k * k
        shape = f"x [{n_},{h_},{w_},{cin}] -> y [{n_},{oh},{ow},{cout}], {k}x{k} s{stride} p{pad}"
        tensor_row(f"da_conv_{tag}_fwd", time_it(lambda: check(lib.da_conv_forward(ctypes.byref(cd), P(cx), P(cw), P(sc), P(sc), 1, 0.0, 0,
                                                                                    P(cy), P(cws), cws.numel(), S()), "conv_forward"), 6), cfl, shape)
        tensor_row(f"da_conv_{tag}_dgrad", time_it(lambda: check(lib.da_conv_backward_data(ctypes.byref(cd), P(cdz), P(cw), -1.0, P(cdx), P(cws),
                                                                                            cws.numel(), S()), "conv_backward_data"), 6), cfl,
                   shape + "; GRL weight -1 folded into the epilogue")
        tensor_row(f"da_conv_{tag}_wgrad", time_it(lambda: check(lib.da_conv_backward_weight(ctypes.byref(cd), P(cx), P(cdz), P(cdw), P(cws),
                                                                                              cws.numel(), S()), "conv_backward_weight"), 6), cfl, shape)

    conv_rows("h1_1x1_c5", 2, H, W, C, 512, 1, 1, 0)
    conv_rows("srm_3x3p3_c5", 2, H + 2, W + 2, C // 4, 9 * C // 4, 3, 1, 3)
    conv_rows("global_3x3s2_c5", 2, H, W, C, C // 2, 3, 2, 1)
    # ---- BASELINE config 4 (instance-level stress): RoIAlign 7x7 of 2048 RoIs over 4 images
    N4, R4 = 4, 4 * ROIS_PER_IMG
    feat4 = torch.relu(torch.randn(N4, H, W, C, device=dev, generator=g)).to(act).permute(0, 3, 1, 2)
    u4 = torch.rand(R4, 4, generator=cpu_g)
    x14, y14 = u4[:, 0] * (W * STRIDE - 33), u4[:, 1] * (H * STRIDE - 33)
    wh4 = torch.exp(torch.log(torch.tensor(16.0)) + u4[:, 2:] * (torch.log(torch.tensor(512.0)) - torch.log(torch.tensor(16.0))))
    rois4 = torch.stack([(torch.arange(R4) // ROIS_PER_IMG).float(), x14, y14, torch.clamp(x14 + wh4[:, 0], max=W * STRIDE),
                         torch.clamp(y14 + wh4[:, 1], max=H * STRIDE)], 1).to(dev)
    b4 = R4 * C * 49 * es + N4 * C * H * W * es + R4 * 20
    lay4 = roi_layout
    hbm_row("roi_align_fwd_config4", time_it(lambda: F_.roi_align(feat4, rois4, 7, 1.0 / STRIDE, out_layout=lay4), 6), b4, "4 images, 2048 RoIs (BASELINE config 4)")
    fr4 = feat4.detach().requires_grad_(True)
    o4 = F_.roi_align(fr4, rois4, 7, 1.0 / STRIDE, out_layout=lay4)
    go4 = torch.randn(o4.shape, device=dev, generator=g).to(o4.dtype)
    hbm_row("roi_align_bwd_config4", time_it(lambda: torch.autograd.grad(o4, fr4, go4, retain_graph=True), 6), b4, "4 images, 2048 RoIs (BASELINE config 4)")
    del feat4, fr4, o4, go4
    # fused SGD + bf16 shadow refresh over the FC1 weight: read w, grad, momentum; write w, momentum, shadow
    gradf = dw.view(-1)
    hbm_row("sgd_step_fc1", time_it(lambda: check(lib.da_sgd_step(P(wf), P(gradf), P(buf), n, 0.01, 0.9, 1e-4, 0, P(shadow), S()), "sgd_step")),
            n * (4 * 3 + 4 * 2 + 2))
    # "next" rows of SURVEY 8f, measured beside the step (not launched by it): the query-axis softmax kernels of the blocked
    # NonLocalBlock on one [32768 x 4096] key block (C3 of a 1024x2048 input), the blocked attention as a whole, and the RPN
    # proposal stage of one image at the train configuration (122880 candidates -> 12000 ranked -> NMS -> 2000)
    del wf, gradf, buf, shadow, dw
    torch.cuda.empty_cache()
    Tq, Tk, Inl = 32768, 4096, 256
    s_blk = torch.randn(Tq, Tk, device=dev, generator=g) * 3
    p_blk = torch.empty(Tq, Tk, dtype=torch.bfloat16, device=dev)
    stats = torch.empty(2 * Tk, device=dev)
    cws = torch.empty(lib.da_colsoftmax_workspace_bytes(Tq, Tk), dtype=torch.uint8, device=dev)
    hbm_row("nlb_colsoftmax_fwd", time_it(lambda: check(lib.da_colsoftmax_forward(P(s_blk), Tq, Tk, Tk, P(p_blk), 1, P(stats), 0, P(cws), cws.numel(), S()),
                                                   "colsoftmax_forward"), 6), Tq * Tk * 10, "key block [32768 x 4096]: S fp32 read twice, P bf16 written")
    dp_blk = torch.randn(Tq, Tk, device=dev, generator=g)
    ds_blk = torch.empty_like(p_blk)
    hbm_row("nlb_colsoftmax_bwd", time_it(lambda: check(lib.da_colsoftmax_backward(P(p_blk), 1, P(dp_blk), Tq, Tk, Tk, P(ds_blk), 1, P(cws), cws.numel(), S()),
                                                   "colsoftmax_backward"), 6), Tq * Tk * 14, "P bf16 + dP fp32 read twice, dS bf16 written")
    del s_blk, p_blk, dp_blk, ds_blk
    mk = lambda sc: (torch.randn(Tq, Inl, device=dev, generator=g) * sc).to(torch.bfloat16)
    th, ph, gg = mk(0.2), mk(0.2), mk(1.0)
    with torch.no_grad():
        t_nlb = time_it(lambda: F_.nonlocal_attention_blocked(th, ph, gg, Tk), 4)
    tensor_row("nlb_blocked_attention_fwd", t_nlb, 4.0 * Tq * Tq * Inl, "T = 32768 tokens, I = 256, 8 key blocks: 2 GEMMs + the softmax passes (HBM-bound on the fp32 score blocks)")
    del th, ph, gg
    A_r, H_r, W_r = 15, 64, 128
    cls_r = torch.randn(A_r, H_r, W_r, device=dev, generator=g) * 2
    reg_r = torch.randn(4 * A_r, H_r, W_r, device=dev, generator=g) * 0.5
    from unsupervised_domain_adaptation_object_detection_implementation_b200.detection import AnchorGenerator
    base_r = AnchorGenerator(strides=[16], ratios=[0.5, 1.0, 2.0], scales=[2, 4, 8, 16, 32]).base[0]
    t_rpn = time_it(lambda: F_.rpn_proposals(cls_r, reg_r, base_r, 16, (1024, 2048), 12000, 2000, 0.7, 0.0), 6)
    table["rpn_proposals_per_image"] = {"ms": round(t_rpn, 4), "note": "122880 candidates -> 12000 ranked -> NMS 0.7 -> 2000; 4 of our launches + torch.sort; "
                                        "latency-sized (one-CTA scan), no roofline row"}
    # the dominant kernel is picked among the kernels the timed step actually launches
    not_in_step = {"fc1_wgrad", "sgd_step_fc1"} if fused_step else {"fc1_wgrad_sgd", "sgd_step_fc1"}
    for k in table:
        table[k]["in_step"] = (k not in not_in_step and not k.startswith("da_conv_") and not k.endswith("_config4")
                               and not k.startswith("nlb_") and not k.startswith("rpn_")
                               and not k.endswith("_rhwc") and not k.endswith("_rchw"))
    table["da_conv_h1_1x1_c5_fwd"]["in_step"] = table["da_conv_h1_1x1_c5_dgrad"]["in_step"] = table["da_conv_h1_1x1_c5_wgrad"]["in_step"] = True
    dom = max((k for k in table if table[k]["in_step"]), key=lambda k: table[k]["ms"])
    d = table[dom]
    # DRAM bytes per launch of the same kernel at the same size, measured once with `ncu --set full` (profiles/)
    traffic = None
    try:
        # round-2 capture (tools/prof_r02.py under ncu --set full, summarised by tools/ncu_summary.py): launches in program order
        launches = json.load(open(os.path.join(ROOT, "profiles", "r02_hot_kernels_ncu.json")))["launches"]
        pick = {"roi_align_fwd": ("roi_align_fwd_tc_kernel", 0), "roi_align_bwd": ("roi_align_bwd_tc_kernel", 0),
                "fc1_fwd": ("umma_nt_kernel<512, 2, 1, 0>", 0), "fc1_dgrad": ("umma_nt_kernel<256, 2, 1, 0>", 0),
                "fc1_wgrad_sgd": ("umma_tn_kernel<256, 2, 1>", 0)}
        prof = {}
        for name, (prefix, inst) in pick.items():
            for rec in launches:
                if rec["kernel"].startswith(prefix) and rec["instance"] == inst:
                    prof[name] = rec
        if roi_layout == "rhwc":     # the RoIAlign kernels of the bin-major layout have their own capture (tools/prof_roi2.py)
            try:
                for rec in json.load(open(os.path.join(ROOT, "profiles", "r02_roi_rhwc_ncu.json")))["launches"]:
                    if rec["kernel"].startswith("roi_align_fwd_tc_kernel"):
                        prof["roi_align_fwd"] = rec
                    if rec["kernel"].startswith("roi_align_bwd_tc_kernel"):
                        prof["roi_align_bwd"] = rec
            except (OSError, KeyError, ValueError):
                pass
        for k, rec in prof.items():
            if k in table:
                table[k]["ncu_dram_bytes"] = rec["dram_bytes"]
                table[k]["ncu_l2_to_sm_bytes"] = rec["l2_to_sm_bytes"]
        traffic = prof.get(dom, {}).get("dram_bytes")
    except (OSError, KeyError, ValueError):
        pass
    if "bytes" in d:
        roof = {"kernel": dom, "bound": "hbm", "achieved": d["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": d["frac"],
                "traffic": traffic, "peak_source": pk["source"], "share_of_step": round(d["ms"] / step_ms, 4)}
    else:
        roof = {"kernel": dom, "bound": "tensor", "achieved": d["tflops"], "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": d["frac"], "traffic": traffic, "peak_source": pk["source"], "share_of_step": round(d["ms"] / step_ms, 4)}
    return roof, table


# ----------------------------------------------------------------------------------------------
# auxiliary workload: MAF image-level heads (BASELINE config 3 shapes, one pair per GPU)
# ----------------------------------------------------------------------------------------------
def run_maf(args, kind="maf"):
    """kind == "maf": SRM heads (mmdet/models/backbones/resnet_da.py:83-118) on C3/C4/C5 + CE on sigmoid (L3) + backward into the
    features (reversed gradient) + SGD.  FLOPs: 2*M*N*K per implicit GEMM with the enlarged extents of Q12
    (1x1 pad 1 -> (H+2)x(W+2); 3x3 pad 3 on that -> (H+6)x(W+6)), forward + data gradient + weight gradient."""
    import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda
    from unsupervised_domain_adaptation_object_detection_implementation_b200 import dist as ddist, functional as F_, hotpath, optim, _lib
    import torch.distributed as dist
    rank, local, world = ddist.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    uda.set_engine(args.engine)
    act = F_.act_dtype()
    # kind == "fpn" (BASELINE config 2b, an extension: no DA config of the reference has a neck): hotpath.FPNHotPath on P2..P5 of a
    # 1024x2048 pair (256 channels, strides 4..32): one image-level head + L1 per level, multi-level RoIAlign of 2x512 RoIs with
    # the device-side level partition, shared FCs, instance head (chain kernel), consistency, backward into every level, SGD.
    fpn = kind == "fpn"
    shapes = [(256, 256, 512), (256, 128, 256), (256, 64, 128), (256, 32, 64)] if fpn else [(512, 128, 256), (1024, 64, 128), (2048, 64, 128)]
    torch.manual_seed(0)
    if fpn:
        model = hotpath.FPNHotPath(256, (4, 8, 16, 32), 1024).to(dev).train()
        gb = torch.Generator().manual_seed(10 + rank)      # boxes as in make_host_inputs: log-uniform 16..512 px on 1024x2048
        u = torch.rand(2, 512, 4, generator=gb)
        bx1, by1 = u[..., 0] * (2048 - 33), u[..., 1] * (1024 - 33)
        lo, hi = torch.log(torch.tensor(16.0)), torch.log(torch.tensor(512.0))
        bw, bh = torch.exp(lo + u[..., 2] * (hi - lo)), torch.exp(lo + u[..., 3] * (hi - lo))
        bxs = torch.stack([bx1, by1, torch.clamp(bx1 + bw, max=2048.0), torch.clamp(by1 + bh, max=1024.0)], -1)
        props = [bxs[i].contiguous().to(dev) for i in range(2)]
    else:
        model = hotpath.MAFHotPath(tuple(c for c, _, _ in shapes)).to(dev).train()
    params = ddist.trainable_parameters(model, [])
    opt = optim.FusedSGD(params, lr=1e-3, momentum=0.9, weight_decay=5e-4)
    reducer = ddist.OverlappedGradAllReduce(params) if world > 1 else None
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    feats = [[torch.relu(torch.randn(2, h, w, c, device=dev, generator=g)).to(act) for c, h, w in shapes] for _ in range(2)]
    flops = 0.0
    if fpn:       # image heads (C -> 512 1x1 per level) + FC 12544 -> 1024 on 1024 RoIs; forward + data + weight gradient
        for c, h, w in shapes:
            flops += 2.0 * (2 * h * w) * c * 512
        flops += 2.0 * 1024 * (256 * 49) * 1024
    else:
        for c, h, w in shapes:
            flops += 2.0 * (2 * (h + 2) * (w + 2)) * c * (c // 4) + 2.0 * (2 * (h + 6) * (w + 6)) * (9 * c // 4) * (9 * c // 4)
    flops *= 3.0

    def step(slot):
        F_.bump_dropout_counter(dev)
        xs = [t.permute(0, 3, 1, 2).requires_grad_(True) for t in feats[slot]]
        if fpn:
            loss, _ = hotpath.parse_losses(model.forward_train(xs, props, [0, 1]))
        else:
            loss, _ = hotpath.parse_losses(model.forward_train(xs[0], xs[1], xs[2], [0, 1]))
        loss.backward()
        if reducer is not None:
            reducer()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss.detach()

    F_.dropout_counter(dev)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            step(i % 2)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphs, note = [], "cuda_graph"
    try:
        if args.no_graph:
            raise RuntimeError("--no-graph")
        pool = None
        for slot in range(2):
            gph = torch.cuda.CUDAGraph()
            _lib.reset_launch_count()
            with torch.cuda.graph(gph, pool=pool):
                step(slot)
            pool = gph.pool()
            graphs.append(gph)
        per_replay = _lib.launch_count()
    except Exception as e:
        graphs, note, per_replay = [], f"eager ({type(e).__name__}: {str(e)[:80]})", 0
        torch.cuda.synchronize()

    def run(i):
        if graphs:
            graphs[i % 2].replay()
        else:
            step(i % 2)

    for i in range(max(args.warmup, 3)):
        run(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    _lib.reset_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        run(i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = ddist.max_over_ranks(e0.elapsed_time(e1), dev)
    clocks = sampler.stop()
    launches = per_replay * args.steps if graphs else _lib.launch_count()
    pk = peaks()
    tf = flops / (ms / args.steps) / 1e9
    if rank == 0:
        print(json.dumps({
            "metric": "da_train_step_img_pairs_per_s", "value": round(world * args.steps / (ms / 1e3), 3), "unit": "img-pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.engine == "umma_bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": "daf_fpn_p2p5_heads_multilevel_roialign_1024x2048" if fpn else "maf_r50dc5_srm_heads_c3c4c5_1024x2048",
                       "pairs_per_gpu": 1, "features": [[2, c, h, w] for c, h, w in shapes],
                       "engine": args.engine, "launch": note,
                       "step": ("image head + L1 on P2..P5, multi-level RoIAlign (2x512 RoIs, device-side level partition), shared FCs, "
                                "instance head + L4 (chain kernel), L7, backward into every level, SGD") if fpn else
                               "SRM x3 + CE-on-sigmoid, backward into the features, SGD",
                       "l2_policy": "activations exceed L2; two input sets alternated", "auxiliary": True},
            "e2e": None, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"kernel": "whole step (GEMM FLOPs only; the FPN step is dominated by memory-bound head tails and RoIAlign)" if fpn
                         else "whole step (implicit GEMMs of the three SRM heads)", "bound": "tensor", "achieved": round(tf, 1),
                         "peak": pk["bf16_tflops_sustained"] or pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": round(tf / (pk["bf16_tflops_sustained"] or pk["bf16_tflops"]), 4), "traffic": None,
                         "peak_source": pk["source"] + " (sustained: kernels timed inside a long step)", "flops_per_step": flops}}))
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


# ----------------------------------------------------------------------------------------------
# CPU reference arm (oracle port; the one place bench.py may execute oracle/).  Nothing of the product package is
# imported here: parameter shapes are the reference's, written out literally.
# ----------------------------------------------------------------------------------------------
def cpu_reference_state():
    """Parameters of the DAF-Org hot path with the reference's shapes and initialisation
    (ImgAlignmentHead resnet_da_daf_org.py:120-146, Shared2FC convfc_bbox_head.py:198-237, InstanceAlignmentHead
    instance_da.py:42-101) + one synthetic pair (seed 0, same generator recipe as the GPU arm) + torch.optim.SGD with the
    reference recipe (faster_rcnn_r50_daf_c2f.py:8)."""
    g = torch.Generator().manual_seed(0)
    st = {"c5": torch.relu(torch.randn(2, C, H, W, generator=g))}
    u = torch.rand(2 * ROIS_PER_IMG, 4, generator=g)
    x1, y1 = u[:, 0] * (W * STRIDE - 33), u[:, 1] * (H * STRIDE - 33)
    wh = torch.exp(torch.log(torch.tensor(16.0)) + u[:, 2:] * (torch.log(torch.tensor(512.0)) - torch.log(torch.tensor(16.0))))
    st["rois"] = torch.stack([(torch.arange(len(u)) >= ROIS_PER_IMG).float(), x1, y1, torch.clamp(x1 + wh[:, 0], max=W * STRIDE),
                              torch.clamp(y1 + wh[:, 1], max=H * STRIDE)], 1)

    def p(*shape, std=0.01):
        return (torch.randn(*shape, generator=g) * std).requires_grad_(True)

    st["sd_img"] = {"conv1.weight": p(512, C, 1, 1, std=0.001), "conv1.bias": p(512, std=0.0), "conv2.weight": p(1, 512, 1, 1, std=0.001),
                    "conv2.bias": p(1, std=0.0)}
    st["sd_ins"] = {"nlb.conv_phi.weight": p(512, 1024, 1, 1), "nlb.conv_theta.weight": p(512, 1024, 1, 1),
                    "nlb.conv_g.weight": p(512, 1024, 1, 1), "nlb.conv_mask.weight": p(1024, 512, 1, 1),
                    "fc1.weight": p(512, 1024), "fc1.bias": p(512, std=0.0), "fc2.weight": p(512, 512), "fc2.bias": p(512, std=0.0),
                    "fc3.weight": p(2, 512, std=0.05), "fc3.bias": p(2, std=0.0)}
    st["fc1_w"] = p(FC_OUT, C * 49, std=(C * 49) ** -0.5)
    st["fc1_b"] = p(FC_OUT, std=0.0)
    st["fc2_w"] = p(FC_OUT, FC_OUT, std=FC_OUT ** -0.5)
    st["fc2_b"] = p(FC_OUT, std=0.0)
    params = list(st["sd_img"].values()) + list(st["sd_ins"].values()) + [st["fc1_w"], st["fc1_b"], st["fc2_w"], st["fc2_b"]]
    st["opt"] = torch.optim.SGD(params, lr=1e-3, momentum=0.9, weight_decay=5e-4)
    return st


def cpu_reference_step(state, threads):
    """ONE FULL train step of the hot path on one source+target pair on the CPU, fp32, no sampling and no scaling:
    H1+L1 on the whole C5, RoIAlign 7x7 of all 2x512 RoIs (oracle/roi_align_ref.c, RoIs split over `threads` host
    threads), shared FCs, I1+L4, L7, backward of all of it (RoIAlign's transposed map in fp32 `+=` arithmetic, channel
    planes split over the threads), SGD with momentum on every parameter.  Returns seconds per component."""
    from oracle import da_oracle, roi_align as oracle_roi
    import numpy as np
    import torch.nn.functional as F
    t, R = {}, 2 * ROIS_PER_IMG
    c5, rois = state["c5"], state["rois"]
    labels = (torch.arange(R) >= ROIS_PER_IMG).long()
    state["opt"].zero_grad(set_to_none=True)
    t0 = time.perf_counter()
    x = c5.clone().requires_grad_(True)
    img_feat = da_oracle.img_alignment_head(x, state["sd_img"])
    g_loss = 0.1 * da_oracle.daf_image_loss(img_feat, torch.tensor([0, 1]))
    t["img_head_fwd"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    pooled, _, _ = oracle_roi.roi_align_forward(c5.numpy(), rois.numpy(), 7, 1.0 / STRIDE, threads=threads)
    t["roi_fwd"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    feats_in = torch.from_numpy(pooled).flatten(1).requires_grad_(True)
    f = F.relu(F.linear(feats_in, state["fc1_w"], state["fc1_b"]))
    f = F.relu(F.linear(f, state["fc2_w"], state["fc2_b"]))
    pred = torch.sigmoid(da_oracle.instance_alignment_logits(f, state["sd_ins"]))
    loss = g_loss + 0.1 * da_oracle.ce2(pred, labels) + 0.1 * da_oracle.consistency_loss(img_feat, pred, labels)
    loss.backward()
    t["fc_instance_fwd_and_all_bwd"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    gin = oracle_roi.roi_align_backward(feats_in.grad.view(R, C, 7, 7).numpy(), rois.numpy(), (2, C, H, W), 7, 1.0 / STRIDE,
                                        threads=threads, dtype=np.float32)
    x.grad.add_(torch.from_numpy(gin))
    t["roi_bwd"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    state["opt"].step()
    t["sgd"] = time.perf_counter() - t0
    return t


def cpu_reference(budget_s):
    """cpu_baseline of the `ours` line: full pairs, as many as fit the budget (at least one after one warm-up)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state = cpu_reference_state()
    cpu_reference_step(state, cores)                       # warm-up (allocator, thread pools, momentum buffers)
    secs, t_all, t0 = [], {}, time.perf_counter()
    while not secs or (time.perf_counter() - t0 + secs[-1] < budget_s and len(secs) < 5):
        t = cpu_reference_step(state, cores)
        secs.append(sum(t.values()))
        t_all = t
    per = sorted(secs)[len(secs) // 2]
    return {"value": round(1.0 / per, 5), "unit": "img-pairs/s", "cores": cores, "kind": "port",
            "sample": f"{len(secs)} full train step(s) of one source+target pair after 1 warm-up (no sub-sampling, no scaling): H1+L1 on all of "
                      f"C5 [2,{C},{H},{W}], RoIAlign fwd/bwd of 2x{ROIS_PER_IMG} RoIs, shared FCs, I1+L4, L7, backward, SGD+momentum; fp32; "
                      f"oracle port (torch CPU on {cores} threads + oracle/roi_align_ref.c on {cores} threads); median {per:.2f} s/step",
            "seconds": {k: round(x, 3) for k, x in t_all.items()}}


def run_reference(args):
    """--impl reference: the CPU implementation of the path, full pairs, `steps` timed after `warmup` untimed."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    state = cpu_reference_state()
    for _ in range(args.warmup):
        cpu_reference_step(state, cores)
    t0 = time.perf_counter()
    comp = {}
    for _ in range(args.steps):
        for k, x in cpu_reference_step(state, cores).items():
            comp[k] = comp.get(k, 0.0) + x
    wall = time.perf_counter() - t0
    v = args.steps / wall
    sample = (f"each step = ONE FULL source+target pair (no sub-sampling, no scaling): H1+L1 on C5 [2,{C},{H},{W}], RoIAlign fwd/bwd of "
              f"2x{ROIS_PER_IMG} RoIs, shared FCs, I1+L4, L7, backward, SGD+momentum; fp32; oracle port (torch CPU + oracle/roi_align_ref.c, "
              f"both on {cores} host threads; the reference's own classes need /root/reference, which does not exist on the GPU box)")
    print(json.dumps({
        "impl": "reference", "metric": "da_train_step_img_pairs_per_s", "value": round(v, 5), "unit": "img-pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * wall / args.steps, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "pairs_per_gpu": 1, "c5": [2, C, H, W], "rois_per_img": ROIS_PER_IMG},
        "cpu_baseline": {"value": round(v, 5), "unit": "img-pairs/s", "cores": cores, "kind": "port", "sample": sample,
                         "seconds_per_step": {k: round(x / args.steps, 3) for k, x in comp.items()}},
        "e2e": {"value": round(v, 5), "unit": "img-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(wall, 1)}))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "fpn":
        run_maf(a, "fpn")
    elif a.workload == "maf":
        run_maf(a)
    else:
        run_ours(a)
