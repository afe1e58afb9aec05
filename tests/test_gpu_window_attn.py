"""tcgen05 shifted-window attention core (occ_window_attention) vs the oracle's ShiftWindowMSA restatement
(port.shift_window_msa with an identity output projection), on token-ordered qkv rows in the S32 split format.  Split-bf16
operands and probabilities (three tensor-core passes): expected error ~1e-5; gate 1e-4 on both SURVEY criteria."""
import pytest
import torch

from oracle import port
from util import assert_close

pytestmark = pytest.mark.gpu


def _rows_from_images(img, B, X, Y, Z):
    """img (B*(Z+1), X, Y, F): images (b,z) then the B BEV images -> token rows (voxel tokens, then BEV tokens)."""
    F_ = img.shape[-1]
    vox = img[:B * Z].view(B, Z, X, Y, F_).permute(0, 2, 3, 1, 4).reshape(B * X * Y * Z, F_)
    bev = img[B * Z:].reshape(B * X * Y, F_)
    return torch.cat([vox, bev], 0).contiguous()


@pytest.mark.parametrize("B,X,Y,Z,C,shift", [(1, 14, 7, 1, 32, False), (1, 10, 16, 2, 128, True), (2, 15, 10, 4, 128, False),
                                             (1, 9, 16, 2, 256, True), (1, 50, 50, 8, 128, True), (1, 7, 7, 1, 1024, True),
                                             (1, 13, 13, 2, 512, False)])
def test_window_attention_vs_oracle(cuda, B, X, Y, Z, C, shift):
    from occformer_b200 import ops
    heads = C // 32
    g = torch.Generator().manual_seed(B * 1000 + X * 10 + C)
    nimg = B * (Z + 1)
    # the qkv projection is three stacked identities, so q = k = v = x + b exactly on both sides and the test isolates the
    # attention core (scores, bias, shift mask, softmax, PV)
    x = torch.randn(nimg, X * Y, C, generator=g)
    sd = {"w_msa.qkv.weight": torch.cat([torch.eye(C)] * 3, 0),
          "w_msa.qkv.bias": 0.3 * torch.randn(3 * C, generator=g),
          "w_msa.proj.weight": torch.eye(C), "w_msa.proj.bias": torch.zeros(C),
          "w_msa.relative_position_bias_table": torch.randn(169, heads, generator=g)}
    qkv_img = torch.nn.functional.linear(x, sd["w_msa.qkv.weight"], sd["w_msa.qkv.bias"]).view(nimg, X, Y, 3 * C)
    ref = port.shift_window_msa(x, (X, Y), sd, "", heads, 3 if shift else 0).view(nimg, X, Y, C)
    qkv_rows = _rows_from_images(qkv_img, B, X, Y, Z)
    table = sd["w_msa.relative_position_bias_table"]
    dense = table[port.rel_position_index(7).view(-1)].view(49, 49, heads).permute(2, 0, 1).reshape(heads, -1)
    bias_pad = torch.nn.functional.pad(dense, (0, 2404 - 2401)).contiguous()
    S = lambda t: ops.to_split(t.contiguous().to(cuda))  # noqa: E731
    bias_s = lambda b: ops.split_weight(b.view(1, -1)).view(-1).to(cuda)  # noqa: E731
    out_s = ops.window_attention(S(qkv_rows), bias_s(sd["w_msa.qkv.bias"]), bias_pad.to(cuda), B, X, Y, Z, C, heads, shift)
    out = ops.from_split(out_s)
    ref_rows = _rows_from_images(ref, B, X, Y, Z)
    assert_close(out, ref_rows, 1e-4, f"window attention B{B} {X}x{Y}x{Z} C{C} shift={shift}")
    # head-major qkv columns ([head][q|k|v][32], what the encoder feeds): identical arithmetic, identical result
    perm = ops.qkv_head_major_perm(C, heads)
    out_hm = ops.window_attention(S(qkv_rows[:, perm]), bias_s(sd["w_msa.qkv.bias"][perm].contiguous()),
                                  bias_pad.to(cuda), B, X, Y, Z, C, heads, shift, head_major=True)
    assert torch.equal(out_hm, out_s), "head-major qkv layout changes the result"


@pytest.mark.parametrize("B,X,Y,Z,shift", [(1, 14, 7, 1, False), (1, 10, 16, 2, True), (2, 15, 10, 4, False),
                                           (1, 50, 50, 8, True), (1, 33, 40, 3, True)])
def test_fused_qkv_window_attention_vs_oracle(cuda, B, X, Y, Z, shift):
    """occ_swin_qkv_attention (QKV projection inside the attention kernel, C = 128) vs the oracle's ShiftWindowMSA with the
    same random qkv weights (identity output projection), and vs the two-kernel path (GEMM + occ_window_attention)."""
    from occformer_b200 import ops
    C, heads = 128, 4
    g = torch.Generator().manual_seed(B * 1000 + X * 10 + Z)
    nimg = B * (Z + 1)
    x = torch.randn(nimg, X * Y, C, generator=g)
    sd = {"w_msa.qkv.weight": torch.randn(3 * C, C, generator=g) * C ** -0.5,
          "w_msa.qkv.bias": 0.3 * torch.randn(3 * C, generator=g),
          "w_msa.proj.weight": torch.eye(C), "w_msa.proj.bias": torch.zeros(C),
          "w_msa.relative_position_bias_table": torch.randn(169, heads, generator=g)}
    ref = port.shift_window_msa(x, (X, Y), sd, "", heads, 3 if shift else 0).view(nimg, X, Y, C)
    rows = _rows_from_images(x.view(nimg, X, Y, C), B, X, Y, Z)
    table = sd["w_msa.relative_position_bias_table"]
    dense = table[port.rel_position_index(7).view(-1)].view(49, 49, heads).permute(2, 0, 1).reshape(heads, -1)
    bias_pad = torch.nn.functional.pad(dense, (0, 2404 - 2401)).contiguous().to(cuda)
    perm = ops.qkv_head_major_perm(C, heads)
    w_hm = ops.split_weight(sd["w_msa.qkv.weight"][perm]).to(cuda)
    b_hm = sd["w_msa.qkv.bias"][perm].contiguous().to(cuda)
    tokn = ops.to_split(rows.to(cuda))
    # the kernel's operand is the window layout of this shift; to_window_layout is host index arithmetic, independent of
    # the device-side mapping in window_geom.cuh (which the encoder tests exercise through occ_gn_relu_zmean_ln)
    out_s = ops.swin_qkv_attention(ops.to_window_layout(tokn, B, X, Y, Z, shift), w_hm, b_hm, bias_pad, B, X, Y, Z, C, heads, shift)
    ref_rows = _rows_from_images(ref, B, X, Y, Z)
    assert_close(ops.from_split(out_s), ref_rows, 1e-4, f"fused qkv+window attention B{B} {X}x{Y}x{Z} shift={shift}")
    qkv = ops.gemm(tokn, w_hm, bias=b_hm, split_out=True)
    two = ops.window_attention(qkv, ops.split_weight(b_hm.cpu().view(1, -1)).view(-1).to(cuda), bias_pad, B, X, Y, Z, C, heads,
                               shift, head_major=True)
    assert_close(ops.from_split(out_s), ops.from_split(two), 5e-5, "fused vs GEMM + window attention")
