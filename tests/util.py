import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star tolerance: 1e-3 relative fp32 per output tensor
RTOL = 1e-3


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_err(a, b):
    """max|a-b| / max|b| (per-tensor relative error)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_close(a, b, tol=RTOL, what=""):
    """per-tensor relative error <= tol AND allclose(rtol=tol, atol=tol*rms(b))."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    r = rel_err(a, b)
    rms = float(b.pow(2).mean().sqrt())
    ok = torch.allclose(a, b, rtol=tol, atol=tol * max(rms, 1e-30))
    frac_bad = float(((a - b).abs() > tol * rms + tol * b.abs()).double().mean())
    assert r <= tol and ok, f"{what}: rel {r:.3e} (tol {tol:.0e}), allclose={ok}, frac_bad={frac_bad:.2e}, rms={rms:.3e}"
    return r


def round_tf32(t):
    i = t.clone().contiguous().view(torch.int32)
    i.add_(0x1000).bitwise_and_(-8192)
    return i.view(torch.float32)
