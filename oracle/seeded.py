"""ORACLE — test infrastructure only; never imported by the product path.

Deterministic, name-keyed parameter / input generation shared by the golden-vector generator
(run against the reference's own classes) and the tests (run against this repo's modules).
Both sides expose identical state_dict keys, so filling parameters by key gives identical
weights without storing them in the fixtures.  Weights use N(0, 1/sqrt(fan_in)) instead of the
reference's N(0, 0.01) so that head outputs are O(1) and errors are visible.
"""
import zlib

import torch


def _gen(name, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def seeded_tensor(name, shape, seed=0, kind="normal", scale=1.0):
    g = _gen(name, seed)
    if kind == "normal":
        return torch.randn(*shape, generator=g) * scale
    if kind == "uniform":  # U(0.5, 1.5) * scale
        return (torch.rand(*shape, generator=g) + 0.5) * scale
    raise ValueError(kind)


@torch.no_grad()
def fill_state_(module, seed=0, prefix=""):
    """Overwrite every parameter and buffer of `module` from its state_dict key."""
    for key, t in module.state_dict().items():
        name = prefix + key
        if key.endswith("num_batches_tracked"):
            continue
        if key.endswith("running_var"):
            v = seeded_tensor(name, t.shape, seed, "uniform")
        elif key.endswith("running_mean"):
            v = seeded_tensor(name, t.shape, seed, "normal", 0.1)
        elif t.dim() >= 2:
            fan_in = t[0].numel()
            v = seeded_tensor(name, t.shape, seed, "normal", fan_in ** -0.5)
        elif key.endswith("weight"):  # BN gamma
            v = seeded_tensor(name, t.shape, seed, "uniform")
        else:  # biases
            v = seeded_tensor(name, t.shape, seed, "normal", 0.1)
        t.copy_(v.to(t.dtype))
    return module


def feature_map(name, shape, seed=0):
    """Post-ReLU-like backbone features: relu(randn)."""
    return torch.relu(seeded_tensor(name, shape, seed))


def synthetic_rois(num_per_img, n_img, img_h, img_w, seed=0, min_size=16.0, max_size=512.0):
    """x1~U(0,W-33), y1~U(0,H-33), w,h = exp(U(ln min, ln max)) clipped."""
    g = _gen("rois", seed)
    R = num_per_img * n_img
    u = torch.rand(R, 4, generator=g)
    x1 = u[:, 0] * (img_w - 33)
    y1 = u[:, 1] * (img_h - 33)
    lo, hi = torch.log(torch.tensor(min_size)), torch.log(torch.tensor(max_size))
    w = torch.exp(lo + u[:, 2] * (hi - lo))
    h = torch.exp(lo + u[:, 3] * (hi - lo))
    x2 = torch.minimum(x1 + w, torch.tensor(float(img_w)))
    y2 = torch.minimum(y1 + h, torch.tensor(float(img_h)))
    b = torch.arange(R) // num_per_img
    return torch.stack([b.float(), x1, y1, x2, y2], 1).contiguous()


def adversarial_rois(n_img, img_h, img_w):
    """Edge cases the reference kernel defines: partly / fully outside, zero area, inverted, huge,
    sub-pixel, exactly on the border."""
    W, H = float(img_w), float(img_h)
    rows = [
        [0, -40.0, -40.0, 60.0, 60.0],          # partly outside (top-left)
        [0, W - 30, H - 30, W + 90, H + 90],    # partly outside (bottom-right)
        [0, -500.0, -500.0, -300.0, -300.0],    # fully outside
        [0, W + 100, H + 100, W + 300, H + 400],
        [0, 100.0, 100.0, 100.0, 100.0],        # zero area
        [0, 200.0, 200.0, 150.0, 120.0],        # inverted (negative size)
        [0, -W, -H, 2 * W, 2 * H],              # huge
        [0, 33.3, 47.7, 34.1, 48.2],            # sub-pixel
        [0, 0.0, 0.0, W, H],                    # whole image
        [0, 0.0, 0.0, 16.0, 16.0],              # one feature cell at stride 16
        [n_img - 1, W / 2, H / 2, W / 2 + 7.0, H / 2 + 300.0],  # thin & tall
        [n_img - 1, 5.0, H - 1.0, W - 5.0, H],  # on the bottom border
    ]
    return torch.tensor(rows, dtype=torch.float32)
