"""Single-pass bf16 mode (``ops.set_precision("bf16")`` = occ_set_mma_passes(1); BASELINE config 5 names bf16).

The reference has no twin of this mode (its fp16 flag covers the image backbone only), so it is validated
against the fp32 oracle at ITS OWN tolerance, stated here: operands carry 8 mantissa bits, i.e. ~2e-3 relative rms per
contraction and ~1e-2 through a block.  Gate: rel_max <= 3e-2 per tensor.  The tests also prove that the switch is live
(the error is far above the 3-pass error) and that the default mode is back afterwards (the 1e-3 gate holds again)."""
import pytest
import torch

from oracle import port
from occformer_b200 import synth
from util import assert_close, rel_err

pytestmark = pytest.mark.gpu

BF16_TOL = 3e-2


@pytest.fixture()
def bf16(cuda):
    from occformer_b200 import ops
    with ops.precision("bf16"):
        yield ops
    assert ops.set_precision("fp32") == "fp32"  # the context manager restored the default


def test_gemm_single_pass_error_level(cuda):
    from occformer_b200 import ops
    g = torch.Generator().manual_seed(0)
    a, w = torch.randn(1000, 256, generator=g), torch.randn(192, 256, generator=g) * 256 ** -0.5
    want = (a.double() @ w.double().t()).float()
    a_s, w_s = ops.to_split(a.to(cuda)), ops.split_weight(w).to(cuda)
    with ops.precision("bf16"):
        got1 = ops.gemm(a_s, w_s)
    got3 = ops.gemm(a_s, w_s)
    e1, e3 = rel_err(got1, want), rel_err(got3, want)
    print(f"[parity] GEMM 1000x192x256: single-pass bf16 rel_max {e1:.2e}, three passes {e3:.2e}")
    assert 2e-4 < e1 < BF16_TOL, "single-pass mode not active or out of its tolerance"
    assert e3 < 5e-5, "default mode not restored"


@pytest.mark.parametrize("cin,c,stride,shift", [(128, 128, 1, True), (128, 256, 2, False)])
def test_block_bf16_mode(bf16, cuda, cin, c, stride, shift):
    from occformer_b200.encoder import DualpathTransformerBlock
    g = torch.Generator().manual_seed(c + stride)
    sd = port.make_block_state(cin, c, stride, g)
    blk = DualpathTransformerBlock(cin, c, stride=stride, norm_cfg=dict(type="GN", num_groups=32, requires_grad=True),
                                   layer_index=1 if shift else 0)
    blk.load_state_dict(sd, strict=True)
    blk = blk.to(cuda).eval()
    x = synth.encoder_input(1, cin, 50, 50, 8, seed=7)
    out = blk(x.to(cuda))
    ref = port.dualpath_block(x, sd, "", stride, shift)
    e = assert_close(out, ref, BF16_TOL, f"bf16 single-pass block {cin}->{c} s{stride} shift={shift}")
    assert e > 2e-4, "single-pass mode not active"


def test_default_mode_after_bf16(cuda):
    """the same block at the fp32-faithful default right after the bf16 tests: the 1e-3 gate"""
    from occformer_b200 import ops
    from occformer_b200.encoder import DualpathTransformerBlock
    assert ops.set_precision("fp32") == "fp32"
    g = torch.Generator().manual_seed(129)
    sd = port.make_block_state(128, 128, 1, g)
    blk = DualpathTransformerBlock(128, 128, stride=1, norm_cfg=dict(type="GN", num_groups=32, requires_grad=True), layer_index=1)
    blk.load_state_dict(sd, strict=True)
    blk = blk.to(cuda).eval()
    x = synth.encoder_input(1, 128, 50, 50, 8, seed=7)
    assert_close(blk(x.to(cuda)), port.dualpath_block(x, sd, "", 1, True), what="default mode after bf16")
