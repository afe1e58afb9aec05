"""Dual-path encoder (CUDA, through the C ABI) vs the oracle port and the reference-generated goldens.
Tolerance = north_star's 1e-3 relative fp32 (tests/util.py)."""
import pytest
import torch

from oracle import port
from occformer_b200 import synth
from util import assert_close, golden, rel_err

pytestmark = pytest.mark.gpu


def _block(cuda, cin, c, stride, shift, sd):
    from occformer_b200.encoder import DualpathTransformerBlock
    blk = DualpathTransformerBlock(cin, c, stride=stride, norm_cfg=dict(type="GN", num_groups=32, requires_grad=True),
                                   layer_index=1 if shift else 0)
    missing, unexpected = blk.load_state_dict(sd, strict=True)
    return blk.to(cuda).eval()


GOLD = [("block_c128_s1_plain", 128, 128, 1, False, (15, 10, 4), 1),
        ("block_c256_s2_shift", 128, 256, 2, True, (15, 10, 4), 2),
        ("block_c128_s1_shift", 128, 128, 1, True, (9, 16, 2), 3)]


@pytest.mark.parametrize("case", GOLD)
def test_block_vs_reference_golden(cuda, case):
    name, cin, c, stride, shift, grid, seed = case
    g = torch.Generator().manual_seed(seed)
    sd = port.make_block_state(cin, c, stride, g)
    blk = _block(cuda, cin, c, stride, shift, sd)
    x = synth.encoder_input(1, cin, *grid, seed=seed + 100)
    out = blk(x.to(cuda))
    gold = torch.from_numpy(golden(name + ".npz")["out"])
    ref = port.dualpath_block(x, sd, "", stride, shift)
    assert rel_err(ref, gold) < 1e-5, "oracle port drifted from the reference golden"
    assert out.shape == gold.shape
    assert_close(out, gold, what=f"{name} vs reference golden")


@pytest.mark.parametrize("cin,c,stride,shift,grid,B", [(128, 128, 1, False, (50, 50, 8), 1),
                                                       (128, 128, 1, True, (50, 50, 8), 2),
                                                       (128, 256, 2, True, (50, 50, 8), 1),
                                                       (256, 512, 2, False, (25, 25, 4), 1),
                                                       (512, 1024, 2, True, (13, 13, 2), 1)])
def test_block_vs_oracle(cuda, cin, c, stride, shift, grid, B):
    g = torch.Generator().manual_seed(c + stride)
    sd = port.make_block_state(cin, c, stride, g)
    blk = _block(cuda, cin, c, stride, shift, sd)
    x = synth.encoder_input(B, cin, *grid, seed=7)
    out = blk(x.to(cuda))
    ref = port.dualpath_block(x, sd, "", stride, shift)
    assert_close(out, ref, what=f"block {cin}->{c} s{stride} shift={shift} grid={grid}")


def test_encoder_pr1(cuda):
    """BASELINE.json configs[0] topology: 50x50x8 voxels, 2-layer dual-path (block_numbers=[1,1])."""
    from occformer_b200.encoder import OccupancyEncoder
    cfg = dict(in_channels=128, num_stage=2, block_numbers=[1, 1], block_inplanes=[128, 256], block_strides=[1, 2],
               out_indices=(0, 1), norm_cfg=dict(type="GN", num_groups=32, requires_grad=True), with_cp=True)
    sd = port.make_encoder_state(128, [128, 256], [1, 1], [1, 2], seed=0)
    enc = OccupancyEncoder(**cfg)
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(cuda).eval()
    x = synth.encoder_input(1, 128, 50, 50, 8, seed=0)
    outs = enc(x.to(cuda))
    refs = port.occupancy_encoder(x, sd, [1, 1], [1, 2], (0, 1))
    assert len(outs) == len(refs) == 2
    for i, (o, r) in enumerate(zip(outs, refs)):
        assert_close(o, r, what=f"encoder out[{i}]")


def test_encoder_full_topology(cuda):
    """The reference topology (4 stages x 2 blocks, C = 128/256/512/1024; occformer_nusc_r50_256x704.py:87-97)
    on the PR1 grid -- error accumulation through all 8 stacked blocks must stay inside the tolerance."""
    from occformer_b200.encoder import OccupancyEncoder
    planes, nums, strides = [128, 256, 512, 1024], [2, 2, 2, 2], [1, 2, 2, 2]
    cfg = dict(in_channels=128, num_stage=4, block_numbers=nums, block_inplanes=planes, block_strides=strides,
               out_indices=(0, 1, 2, 3), norm_cfg=dict(type="GN", num_groups=32, requires_grad=True), with_cp=True)
    sd = port.make_encoder_state(128, planes, nums, strides, seed=3)
    enc = OccupancyEncoder(**cfg)
    enc.load_state_dict(sd, strict=True)
    enc = enc.to(cuda).eval()
    x = synth.encoder_input(1, 128, 50, 50, 8, seed=11)
    outs = enc(x.to(cuda))
    refs = port.occupancy_encoder(x, sd, nums, strides, (0, 1, 2, 3))
    for i, (o, r) in enumerate(zip(outs, refs)):
        assert o.shape == r.shape
        assert_close(o, r, what=f"full encoder out[{i}] {tuple(r.shape)}")


@pytest.mark.parametrize("M", [1000, 128 * 148 * 3 + 77, 128 * 148 * 2])
def test_fused_swin_tail_vs_fp64(cuda, M):
    """occ_swin_proj_ffn (proj + residual, LN2, FFN, residual in one kernel, C = 128) vs fp64 torch and vs the unfused
    kernels; M covers a single partial tile, odd and even tile counts per CTA."""
    import torch.nn.functional as F
    from occformer_b200 import ops
    C = 128
    g = torch.Generator().manual_seed(M)
    att = torch.randn(M, C, generator=g).to(cuda)
    att_s = ops.to_split(att)
    tok = torch.randn(M, C, generator=g).to(cuda)
    wf = [(torch.randn(C, C, generator=g) * C ** -0.5) for _ in range(3)]
    ws = [ops.split_weight(w).to(cuda) for w in wf]
    bs = [(0.1 * torch.randn(C, generator=g)).to(cuda) for _ in range(3)]
    lw, lb = (1 + 0.1 * torch.randn(C, generator=g)).to(cuda), (0.1 * torch.randn(C, generator=g)).to(cuda)
    out = ops.swin_proj_ffn(att_s, tok, ws[0], bs[0], lw, lb, ws[1], bs[1], ws[2], bs[2])
    y1 = ops.gemm(att_s, ws[0], bias=bs[0], residual=tok)
    y1n = ops.layernorm(y1, lw, lb, split_out=True)
    h = ops.gemm(y1n, ws[1], bias=bs[1], act=2, split_out=True)
    unfused = ops.gemm(h, ws[2], bias=bs[2], residual=y1)
    assert_close(out, unfused, 5e-5, f"fused swin tail vs unfused kernels M={M}")
    d = lambda t: t.double().cpu()
    r1 = d(tok) + d(att) @ d(wf[0]).T + d(bs[0])
    rn = F.layer_norm(r1, (C,), d(lw), d(lb), 1e-5)
    ref = r1 + F.gelu(rn @ d(wf[1]).T + d(bs[1])) @ d(wf[2]).T + d(bs[2])
    assert_close(out, ref, 5e-5, f"fused swin tail vs fp64 M={M}")
