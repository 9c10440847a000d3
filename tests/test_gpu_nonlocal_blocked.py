"""GPU parity of the blocked NonLocalBlock attention (csrc/colsoftmax.cu,
functional.nonlocal_attention_blocked) against the formula of the reference block
(mmdet/models/backbones/resnet_da_deep.py:402-445: softmax(dim=1) of theta^T.phi on [b,T,T] = over the QUERY axis)
evaluated with plain torch in fp64, and against the materialised path of da_heads.NonLocalBlock.

Tolerances: fp32 engines <= 1e-5 of the reference tensor's max magnitude; the bf16 tensor-core engine is stated
separately (bf16 operands and a bf16 P matrix, fp32 accumulation: 3e-2)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

import unsupervised_domain_adaptation_object_detection_implementation_b200 as uda  # noqa: E402
from unsupervised_domain_adaptation_object_detection_implementation_b200 import functional as F_, da_heads, _lib  # noqa: E402
from oracle import seeded  # noqa: E402
from helpers import rel_err  # noqa: E402

DEV = "cuda"
FP32_TOL = 1e-5
BF16_TOL = 3e-2


@pytest.fixture(autouse=True)
def _engine_reset():
    yield
    uda.set_engine("umma_bf16")


def _colsoftmax(s, p_dtype):
    Tq, Tk = s.shape
    p = torch.empty((Tq, Tk), dtype=p_dtype, device=s.device)
    stats = torch.empty((2 * Tk,), dtype=torch.float32, device=s.device)
    ws = torch.empty((F_.lib.da_colsoftmax_workspace_bytes(Tq, Tk),), dtype=torch.uint8, device=s.device)
    F_.check(F_.lib.da_colsoftmax_forward(F_._ptr(s), Tq, Tk, Tk, F_._ptr(p), F_._code(p_dtype), F_._ptr(stats), 0, F_._ptr(ws), ws.numel(),
                                          F_._stream()), "colsoftmax_forward")
    return p, stats, ws


@pytest.mark.parametrize("Tq,Tk", [(1000, 232), (4096, 2048), (257, 33), (8, 2), (70001, 64)])
@pytest.mark.parametrize("p_dtype", [torch.float32, torch.bfloat16])
def test_colsoftmax_kernels_vs_torch(Tq, Tk, p_dtype):
    """Two-pass column softmax and its backward on ragged / odd / tall shapes (vector and scalar code paths)."""
    s = (seeded.seeded_tensor(f"cs.s.{Tq}.{Tk}", (Tq, Tk), 0, scale=4.0)).to(DEV)
    p, stats, ws = _colsoftmax(s, p_dtype)
    ref = torch.softmax(s.double(), dim=0)
    tol = FP32_TOL if p_dtype == torch.float32 else 2.0 ** -8
    assert float((p.double() - ref).abs().max()) <= tol * float(ref.max())
    assert float((p.double().sum(0) - 1).abs().max()) <= (1e-5 if p_dtype == torch.float32 else 4e-3)
    assert torch.allclose(stats[:Tk].double(), s.double().max(0).values)
    assert rel_err(1.0 / stats[Tk:].double(), torch.exp(s.double() - s.double().max(0).values).sum(0)) <= FP32_TOL
    # statistics reused (the backward's recompute path): bit-identical P
    p2 = torch.empty_like(p)
    F_.check(F_.lib.da_colsoftmax_forward(F_._ptr(s), Tq, Tk, Tk, F_._ptr(p2), F_._code(p_dtype), F_._ptr(stats), 1, None, 0, F_._stream()), "fwd")
    assert torch.equal(p, p2)
    # backward: ds = p * (dp - sum_q p dp) with the p the kernel was given
    dp = seeded.seeded_tensor(f"cs.dp.{Tq}.{Tk}", (Tq, Tk), 1).to(DEV)
    for ds_dtype in (torch.float32, p_dtype):
        ds = torch.empty((Tq, Tk), dtype=ds_dtype, device=DEV)
        F_.check(F_.lib.da_colsoftmax_backward(F_._ptr(p), F_._code(p_dtype), F_._ptr(dp), Tq, Tk, Tk, F_._ptr(ds), F_._code(ds_dtype),
                                               F_._ptr(ws), ws.numel(), F_._stream()), "colsoftmax_backward")
        pd = p.double()
        ref_ds = pd * (dp.double() - (pd * dp.double()).sum(0, keepdim=True))
        assert rel_err(ds, ref_ds) <= (FP32_TOL if ds_dtype == torch.float32 else 2.0 ** -8)


def test_colsoftmax_rejects_bad_arguments():
    s = torch.zeros((8, 8), device=DEV)
    rc = F_.lib.da_colsoftmax_forward(F_._ptr(s), 8, 8, 4, F_._ptr(s), _lib.DA_F32, F_._ptr(s), 0, None, 0, F_._stream())
    assert rc != 0 and b"colsoftmax_forward" in F_.lib.da_last_error()


def _attention_ref(theta, phi, g):
    p = torch.softmax(theta @ phi.t(), dim=0)        # over the query axis (rows), per key column
    return p @ g


@pytest.mark.parametrize("engine,tol", [("simt_f32", FP32_TOL), ("umma_bf16x6", FP32_TOL), ("umma_bf16", BF16_TOL)])
@pytest.mark.parametrize("T,I,bk", [(1024, 128, 256), (1000, 64, 384), (512, 256, 2048)])
def test_nonlocal_attention_blocked_vs_fp64_formula(engine, tol, T, I, bk):
    """Forward and the three input gradients of the blocked attention (4 blocks / ragged last block / one block) against the
    reference formula in fp64 on the same (dtype-rounded) inputs."""
    uda.set_engine(engine)
    dt = F_.act_dtype()
    mk = lambda n, sc: (seeded.seeded_tensor(f"nlb.{n}.{T}.{I}", (T, I), 0, scale=sc)).to(DEV).to(dt)
    theta, phi, g = mk("theta", 0.3), mk("phi", 0.3), mk("g", 1.0)
    cot = seeded.seeded_tensor(f"nlb.cot.{T}.{I}", (T, I), 1).to(DEV)
    a = [t.clone().requires_grad_(True) for t in (theta, phi, g)]
    y = F_.nonlocal_attention_blocked(*a, block_k=bk)
    (y.float() * cot).sum().backward()
    b = [t.double().requires_grad_(True) for t in (theta, phi, g)]
    yr = _attention_ref(*b)
    (yr * cot.double()).sum().backward()
    assert rel_err(y, yr) <= tol
    for u, v, name in zip(a, b, ("theta", "phi", "g")):
        assert rel_err(u.grad, v.grad) <= (tol if engine != "umma_bf16" else 6e-2), name


@pytest.mark.parametrize("engine,tol", [("simt_f32", FP32_TOL), ("umma_bf16", BF16_TOL)])
def test_nonlocal_block_blocked_path_matches_materialised_path(engine, tol, monkeypatch):
    """da_heads.NonLocalBlock switches to the blocked attention above BLOCK_TOKENS tokens: same output, same input and
    parameter gradients as its materialised T x T path (which is the one pinned to the reference module's golden fixture)."""
    uda.set_engine(engine)
    torch.manual_seed(0)
    C, T = 256, 768
    blk = da_heads.NonLocalBlock(C).to(DEV)
    for m in (blk.conv_phi, blk.conv_theta, blk.conv_g, blk.conv_mask):
        torch.nn.init.normal_(m.weight, 0, 0.05)
    x = seeded.seeded_tensor("nlb.block.x", (T, C), 0).to(DEV).to(F_.act_dtype())
    cot = seeded.seeded_tensor("nlb.block.cot", (T, C), 1).to(DEV)

    def run():
        xi = x.clone().requires_grad_(True)
        blk.zero_grad(set_to_none=True)
        y = blk.forward_tokens(xi, grl=-1.0)
        (y.float() * cot).sum().backward()
        return y.detach().float(), xi.grad.float(), [p.grad.detach().float().clone() for p in blk.parameters()]

    y0, gx0, gp0 = run()
    monkeypatch.setattr(da_heads.NonLocalBlock, "BLOCK_TOKENS", 256)
    monkeypatch.setattr(da_heads.NonLocalBlock, "BLOCK_K", 256)
    y1, gx1, gp1 = run()
    assert rel_err(y1, y0) <= tol
    assert rel_err(gx1, gx0) <= (tol if engine != "umma_bf16" else 6e-2)
    for u, v in zip(gp1, gp0):
        assert rel_err(u, v) <= (tol if engine != "umma_bf16" else 6e-2)


def test_nonlocal_alignment_head_at_full_c3_size():
    """H5 at the size the materialised form cannot run (T = 128*256 = 32768 tokens per image, A = 4.3 GB in fp32):
    forward + backward of NonLocalAlignmentHead(512) on one C3 map of a 1024x2048 input, and the attention core against a
    chunked fp32 torch evaluation of the same formula on the same bf16 inputs."""
    uda.set_engine("umma_bf16")
    T, I = 128 * 256, 256
    mk = lambda n, sc: seeded.seeded_tensor(f"nlb.full.{n}", (T, I), 0, scale=sc).to(DEV).to(torch.bfloat16)
    theta, phi, g = mk("theta", 0.2), mk("phi", 0.2), mk("g", 1.0)
    torch.cuda.reset_peak_memory_stats()
    y = F_.nonlocal_attention_blocked(theta, phi, g, block_k=2048)
    ref = torch.zeros((T, I), dtype=torch.float32, device=DEV)
    for k0 in range(0, T, 4096):
        p = torch.softmax(theta.float() @ phi[k0:k0 + 4096].float().t(), dim=0)
        ref += p @ g[k0:k0 + 4096].float()
    assert rel_err(y, ref) <= BF16_TOL
    assert torch.cuda.max_memory_allocated() < 6 * (1 << 30)          # nowhere near T*T*4 = 4.3 GB per live matrix (x3 in autograd)
    head = da_heads.NonLocalAlignmentHead(512).to(DEV)
    head.bn1.eval()          # frozen statistics, as in the reference (norm_eval=True)
    x = torch.relu(seeded.seeded_tensor("nlb.full.x", (1, 512, 128, 256), 0)).to(DEV).requires_grad_(True)
    out = head(x)
    assert out.shape == (1, 512, 128, 256) and bool(torch.isfinite(out).all())
    out.float().square().mean().backward()
    assert bool(torch.isfinite(x.grad).all()) and float(x.grad.abs().max()) > 0
    for p_ in head.nlb1.parameters():
        assert p_.grad is not None and bool(torch.isfinite(p_.grad).all())
