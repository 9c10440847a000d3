"""ncu driver: the image-level head (ImgAlignmentHead + L1) at the bench shape, forward + backward, warm."""
import sys, torch
sys.path.insert(0, ".")
import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda
from unsupervised_domain_adaptation_object_detection_implementation_b200 import da_heads
dev = "cuda"
uda.set_engine("umma_bf16")
torch.manual_seed(0)
m = da_heads.ImgAlignmentHead(2048).to(dev).train()
x = torch.relu(torch.randn(2, 64, 128, 2048, device=dev)).to(torch.bfloat16).permute(0, 3, 1, 2).requires_grad_(True)
dom = torch.tensor([0, 1], device=dev, dtype=torch.int32)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.zero_()
    loss, feat = m.forward_loss(x, dom)
    flush.zero_()
    loss.backward()
torch.cuda.synchronize()
print("ok")
