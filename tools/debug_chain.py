"""Debug: intermediates of the instance-head chain kernel against the oracle (GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, torch.nn.functional as F
import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda
from oracle import da_oracle, seeded
from helpers import build_head
_q = da_oracle._q
dev = "cuda"
for R in (24, 300):
    m = build_head("instance_alignment", 0)
    x = seeded.feature_map("chain.instance_alignment.x", (R, 1024), 0).bfloat16().float()
    sd = {k: v.detach().double() for k, v in m.state_dict().items()}
    q = "bf16"
    X = x.double()
    th = _q(F.linear(X, _q(sd["nlb.conv_theta.weight"].flatten(1), q)), q)
    ph = _q(F.linear(X, _q(sd["nlb.conv_phi.weight"].flatten(1), q)), q)
    g = _q(F.linear(X, _q(sd["nlb.conv_g.weight"].flatten(1), q)), q)
    S = th @ ph.t()
    P = _q(torch.softmax(S, dim=0), q)
    Y = _q(P @ g, q)
    t = _q(F.linear(Y, _q(sd["nlb.conv_mask.weight"].flatten(1), q)), q) + X
    h1 = _q(F.relu(F.linear(t, _q(sd["fc1.weight"], q), sd["fc1.bias"])), q)
    h2 = _q(F.relu(F.linear(h1, _q(sd["fc2.weight"], q), sd["fc2.bias"])), q)
    z = F.linear(h2, _q(sd["fc3.weight"], q), sd["fc3.bias"])
    m = m.to(dev)
    loss, pred = m.forward_loss(x.to(dev).requires_grad_(True), torch.zeros(R, dtype=torch.long, device=dev))
    sv = loss.grad_fn.keep[2]
    I = 512
    def err(a, b):
        a, b = a.double().cpu(), b.double()
        return f"max {float((a - b).abs().max() / b.abs().max()):.2e} fro {float((a - b).norm() / b.norm()):.2e}"
    print(f"R={R}  |S| max {float(S.abs().max()):.1f}  P max {float(P.max()):.3f}")
    print("  theta", err(sv["proj"][:, :I], th), " phi", err(sv["proj"][:, I:2 * I], ph), " g", err(sv["proj"][:, 2 * I:], g))
    print("  P", err(sv["attn"][:, :R], P), " Y", err(sv["y"], Y), " t", err(sv["t"], t))
    print("  h1", err(sv["h1"], h1), " h2", err(sv["h2"], h2), " z", err(sv["z"], z))
