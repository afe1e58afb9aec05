"""Torch-tensor wrappers over the C-ABI kernels (device memory + streams are torch's; the math is ours).

Every function checks device / dtype / contiguity, passes raw pointers and the current torch stream to
libocc_b200.so and raises RuntimeError on a non-zero return code.  No fallback paths.
"""
import ctypes

import torch

from ._lib import check, lib

LAUNCH_COUNT = [0]  # kernels launched by this library (bench.py reports it as gpu_launches)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _stream(t=None):
    """the current torch stream of the tensor's device (tensors of one call share a device, checked by _chk)"""
    return ctypes.c_void_p(torch.cuda.current_stream(t.device if t is not None else None).cuda_stream)


def _chk(t, name, dtype=torch.float32):
    if not t.is_cuda:
        raise RuntimeError(f"occformer_b200: {name} must be a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise RuntimeError(f"occformer_b200: {name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"occformer_b200: {name} must be contiguous")
    if t.device.index != torch.cuda.current_device():
        raise RuntimeError(f"occformer_b200: {name} lives on {t.device} but the current device is "
                           f"cuda:{torch.cuda.current_device()} (wrap the call in torch.cuda.device(...))")
    return t


# ----------------------------------------------------------------------------------------------- S32 split format
# Operands of every tensor-core contraction are fp32 values stored as bf16 hi | lo halves per 32-column chunk (csrc/
# occ_ptx.cuh, include/occ_b200.h): same bytes and row pitch as the fp32 tensor, ~1e-5 relative error after the three
# bf16 passes.  Weights are split once at load time (torch ops below), activations by the producing kernel's epilogue.
def set_precision(mode):
    """"fp32" (default): every contraction in three bf16 tensor-core passes on hi/lo split operands (fp32-faithful, the
    mode of every parity claim).  "bf16": single pass on the hi halves (BASELINE config 5; ~3e-3 relative error per
    contraction, no reference twin).  Library-wide, takes effect for kernels launched afterwards.  Returns the old mode."""
    passes = {"fp32": 3, "bf16": 1}[mode]
    prev = lib().occ_set_mma_passes(passes)
    if prev < 0:
        raise RuntimeError("occ_set_mma_passes failed")
    return {3: "fp32", 1: "bf16"}[prev]


class precision:
    """``with ops.precision("bf16"): ...`` -- scoped set_precision."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = set_precision(self.mode)
        return self

    def __exit__(self, *exc):
        set_precision(self.prev)
        return False


def split_weight(w):
    """(N, K) fp32 torch tensor (any device), K % 32 == 0 -> the same tensor in S32 (fp32 container, same shape)."""
    w = w.detach().float().contiguous()
    N, K = w.shape
    assert K % 32 == 0, f"S32 rows need K % 32 == 0, got {K}"
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    packed = torch.cat([hi.view(N, K // 32, 32), lo.view(N, K // 32, 32)], dim=2).contiguous()  # (N, K/32, 64) bf16
    return packed.view(torch.float32).reshape(N, K)


def unsplit_weight(ws):
    """inverse of split_weight (tests / debugging): S32 -> fp32 values hi + lo."""
    N, K = ws.shape
    p = ws.contiguous().view(torch.bfloat16).view(N, K // 32, 64).float()
    return (p[:, :, :32] + p[:, :, 32:]).reshape(N, K)


def to_split(x):
    """fp32 CUDA tensor (..., C), C % 32 == 0 -> S32 (device kernel)."""
    _chk(x, "x")
    C = x.shape[-1]
    out = torch.empty_like(x)
    check(lib().occ_split_rows(_ptr(x), _ptr(out), x.numel() // C, C, _stream(x)), "occ_split_rows")
    LAUNCH_COUNT[0] += 1
    return out


def from_split(x):
    _chk(x, "x")
    C = x.shape[-1]
    out = torch.empty_like(x)
    check(lib().occ_unsplit_rows(_ptr(x), _ptr(out), x.numel() // C, C, _stream(x)), "occ_unsplit_rows")
    LAUNCH_COUNT[0] += 1
    return out


# ----------------------------------------------------------------------------------------------- voxel pooling
class VoxelPoolWorkspace:
    """View of the pooling workspace for one call with n_points points (bookkeeping left behind by the kernels)."""

    def __init__(self, buf, n_points, B, X, Y, Z):
        offs = [ctypes.c_size_t() for _ in range(4)]
        lib().occ_voxel_pool_workspace_layout(n_points, B, X, Y, Z, *[ctypes.byref(o) for o in offs])
        self.buf, self.nbytes = buf, buf.numel()
        self.off_counts, self.off_head, self.off_next, self.off_vox_id = (o.value for o in offs)
        self.V = B * X * Y * Z
        self.P = n_points

    def _ints(self, off, n):
        return self.buf[off:off + 4 * n].view(torch.int32)

    @property
    def counts(self):
        """points per voxel (= the reference's interval lengths)"""
        return self._ints(self.off_counts, self.V)

    @property
    def head(self):
        """first point (1-based id, 0 = empty) of every voxel's list"""
        return self._ints(self.off_head, self.V)

    @property
    def next(self):
        """next point (1-based id, 0 = end) of every kept point"""
        return self._ints(self.off_next, self.P)

    @property
    def vox_id(self):
        return self._ints(self.off_vox_id, self.P)


_ws_cache = {}  # (B, X, Y, Z, device) -> (uint8 buffer, point capacity): one grow-only buffer per grid


def _workspace(n_points, B, X, Y, Z, device):
    key = (B, X, Y, Z, str(device))
    n_points = max(int(n_points), 1)
    ent = _ws_cache.get(key)
    if ent is None or ent[1] < n_points:
        cap = 1 << (n_points - 1).bit_length()  # next power of two: the drop-in bev_pool sees a different n every sample
        nbytes = lib().occ_voxel_pool_workspace_bytes(cap, B, X, Y, Z)
        ent = (torch.zeros(nbytes, dtype=torch.uint8, device=device), cap)
        _ws_cache[key] = ent
    return VoxelPoolWorkspace(ent[0], n_points, B, X, Y, Z)


_conv_ws = {}  # (device, stream) -> zeroed buffer: split-K ordering counters of occ_conv_bf16x3


def _conv_workspace(device):
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _conv_ws.get(key)
    if ws is None:
        ws = torch.zeros(lib().occ_conv_workspace_bytes(), dtype=torch.uint8, device=device)
        _conv_ws[key] = ws
    return ws


def lss_geometry(frustum, rots, trans, intrins, post_rots, post_trans, bda):
    """frustum (D,fH,fW,3) + camera matrices -> geom (B,N,D,fH,fW,3)."""
    B, N = trans.shape[:2]
    D, fH, fW, _ = frustum.shape
    f = lambda t: _chk(t.float().contiguous(), "camera matrix")  # noqa: E731
    geom = torch.empty((B, N, D, fH, fW, 3), dtype=torch.float32, device=frustum.device)
    fr, r, t, k, pr, pt, bd = f(frustum), f(rots), f(trans), f(intrins), f(post_rots), f(post_trans), f(bda)
    check(lib().occ_lss_geometry(_ptr(fr), D * fH * fW, _ptr(r), _ptr(t), _ptr(k), k.shape[-2], k.shape[-1], _ptr(pr),
                                 _ptr(pt), _ptr(bd), bd.shape[-1], B, N, _ptr(geom), _stream(geom)), "occ_lss_geometry")
    LAUNCH_COUNT[0] += 1
    return geom


def lift_prologue(depth_logits, img_feat):
    """depth_logits (BN,D,fH,fW), img_feat (BN,C,fH,fW) -> depth_prob (BN,D,fH,fW), feat_cl (BN,fH*fW,C)."""
    _chk(depth_logits, "depth_logits"), _chk(img_feat, "img_feat")
    BN, D, fH, fW = depth_logits.shape
    C = img_feat.shape[1]
    prob = torch.empty_like(depth_logits)
    feat_cl = torch.empty((BN, fH * fW, C), dtype=torch.float32, device=img_feat.device)
    check(lib().occ_lift_prologue(_ptr(depth_logits), _ptr(img_feat), _ptr(prob), _ptr(feat_cl), BN, D, C, fH * fW,
                                  _stream()), "occ_lift_prologue")
    LAUNCH_COUNT[0] += 2
    return prob, feat_cl


def lift_splat(depth_prob, feat_cl, geom, B, N, dx, bx, nx, grid, return_workspace=False, with_split=False):
    """Fused lift-splat.  Returns the channel-last grid (B,X,Y,Z,C) [, its S32 copy when with_split]."""
    _chk(depth_prob, "depth_prob"), _chk(feat_cl, "feat_cl"), _chk(geom, "geom")
    BN, D, fH, fW = depth_prob.shape
    C = feat_cl.shape[-1]
    X, Y, Z = grid
    P = B * N * D * fH * fW
    assert geom.numel() == 3 * P
    ws = _workspace(P, B, X, Y, Z, geom.device)
    out = torch.empty((B, X, Y, Z, C), dtype=torch.float32, device=geom.device)
    out_s = torch.empty_like(out) if with_split else None
    f = [float(v) for v in (*dx, *bx, *nx)]
    check(lib().occ_lift_splat(_ptr(depth_prob), _ptr(feat_cl), _ptr(geom), _ptr(out), _ptr(out_s), B, N, D, fH * fW, C,
                               *f, X, Y, Z, _ptr(ws.buf), ws.nbytes, _stream(geom)), "occ_lift_splat")
    LAUNCH_COUNT[0] += 3
    res = (out, out_s) if with_split else out
    return (res, ws) if return_workspace else res


def lift_splat_fused(depth_logits, img_feat, frustum, rots, trans, intrins, post_rots, post_trans, bda, B, N, dx, bx, nx,
                     grid, with_split=False):
    """The view transformer after DepthNet in three launches (memset, front kernel, pooling): depth_logits (BN,D,fH,fW),
    img_feat (BN,C,fH,fW) -> (grid (B,X,Y,Z,C) [, S32 twin], depth_prob (BN,D,fH,fW))."""
    BN, D, fH, fW = depth_logits.shape
    C = img_feat.shape[1]
    X, Y, Z = grid
    HW = fH * fW

    def plane(t, name):  # (BN, ch, fH, fW) fp32 whose cameras may be strided (channel slice of a bigger tensor): no copy
        t = t.float()
        if t.stride()[1:] != (HW, fW, 1):
            t = t.contiguous()
        if not t.is_cuda:
            raise RuntimeError(f"occformer_b200: {name} must be a CUDA tensor (no CPU fallback exists)")
        return t, t.stride(0)

    depth_logits, ls = plane(depth_logits, "depth_logits")
    img_feat, fs = plane(img_feat, "img_feat")
    f = lambda t: _chk(t.float().contiguous(), "camera matrix")  # noqa: E731
    fr, r, t, k, pr, pt, bd = f(frustum), f(rots), f(trans), f(intrins), f(post_rots), f(post_trans), f(bda)
    dev = depth_logits.device
    ws = _workspace(BN * D * HW, B, X, Y, Z, dev)
    prob = torch.empty((BN, D, fH, fW), dtype=torch.float32, device=dev)
    feat_cl = torch.empty((BN, HW, C), dtype=torch.float32, device=dev)
    out = torch.empty((B, X, Y, Z, C), dtype=torch.float32, device=dev)
    out_s = torch.empty_like(out) if with_split else None
    fl = [float(v) for v in (*dx, *bx, *nx)]
    check(lib().occ_lift_splat_fused(_ptr(depth_logits), ls, _ptr(img_feat), fs, _ptr(fr), _ptr(r), _ptr(t), _ptr(k), k.shape[-2],
                                     k.shape[-1], _ptr(pr), _ptr(pt), _ptr(bd), bd.shape[-1], _ptr(prob), _ptr(feat_cl),
                                     _ptr(out), _ptr(out_s), B, N, D, HW, C, *fl, X, Y, Z, _ptr(ws.buf), ws.nbytes,
                                     _stream(prob)), "occ_lift_splat_fused")
    LAUNCH_COUNT[0] += 3
    return ((out, out_s) if with_split else out), prob


def voxel_pool_geom(feats, geom, B, dx, bx, nx, grid, return_workspace=False):
    """Materialised-volume voxel pooling: feats (P,C), geom (P,3) -> (B,X,Y,Z,C)."""
    _chk(feats, "feats"), _chk(geom, "geom")
    P_, C = feats.shape
    X, Y, Z = grid
    ws = _workspace(P_, B, X, Y, Z, feats.device)
    out = torch.empty((B, X, Y, Z, C), dtype=torch.float32, device=feats.device)
    f = [float(v) for v in (*dx, *bx, *nx)]
    check(lib().occ_voxel_pool_geom(_ptr(feats), _ptr(geom), _ptr(out), B, P_ // B, C, *f, X, Y, Z, _ptr(ws.buf),
                                    ws.nbytes, _stream()), "occ_voxel_pool_geom")
    LAUNCH_COUNT[0] += 3
    return (out, ws) if return_workspace else out


def bev_pool_channel_last(feats, coords, B, X, Y, Z, return_workspace=False):
    _chk(feats, "feats"), _chk(coords, "coords", torch.int64)
    n, C = feats.shape
    ws = _workspace(n, B, X, Y, Z, feats.device)
    out = torch.empty((B, X, Y, Z, C), dtype=torch.float32, device=feats.device)
    check(lib().occ_bev_pool(_ptr(feats), _ptr(coords), _ptr(out), n, C, B, X, Y, Z, _ptr(ws.buf), ws.nbytes,
                             _stream()), "occ_bev_pool")
    LAUNCH_COUNT[0] += 3
    return (out, ws) if return_workspace else out


# ----------------------------------------------------------------------------------------------- GEMM / conv
def gemm(a, w, bias=None, residual=None, act=0, split_out=False, out=None):
    """out[M,N] = act(a[M,K] @ w[N,K]^T + bias) (+ residual).  a, w in S32 (to_split / split_weight / a producer's
    split output); out fp32, or S32 when split_out (operand of the next contraction)."""
    _chk(a, "a"), _chk(w, "w")
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    else:
        _chk(out, "out")
    if bias is not None:
        _chk(bias, "bias")
    if residual is not None:
        _chk(residual, "residual")
        assert residual.shape == (M, N)
    check(lib().occ_gemm_bf16x3(_ptr(a), _ptr(w), _ptr(out), M, N, K, _ptr(bias), _ptr(residual), act, int(split_out),
                                None, 0, 0, _stream(a)), "occ_gemm_bf16x3")
    LAUNCH_COUNT[0] += 1
    return out


def repack_conv_weight(w):
    """(Cout, Cin, kx, ky[, kz]) -> (Cout, taps*Cin) tap-major [(kx*KY + ky)*KZ + kz][cin] in S32 (Cin % 32 == 0)."""
    if w.dim() == 4:
        w = w.unsqueeze(-1)
    Cout, Cin, KX, KY, KZ = w.shape
    w2 = w.detach().float().permute(0, 2, 3, 4, 1).reshape(Cout, KX * KY * KZ * Cin)
    return split_weight(w2), (KX, KY, KZ)


def conv(x_cl, w2, ksize, stride=1, dil=1, bias=None, residual=None, act=0, split_out=False, gn_stats=None, cpg=0):
    """x_cl (B,X,Y,Z,Cin) channel-last S32; w2 from repack_conv_weight; returns (B,Xo,Yo,Zo,Cout) channel-last raw
    output (fp32, or S32 when split_out).  gn_stats: zero-initialised double tensor (B, groups, 2) receiving (sum, sumsq)."""
    _chk(x_cl, "x_cl"), _chk(w2, "w2")
    B, X, Y, Z, Cin = x_cl.shape
    KX, KY, KZ = ksize
    Cout = w2.shape[0]
    assert w2.shape[1] == KX * KY * KZ * Cin

    def osz(n, k):
        p = dil * (k - 1) // 2
        return (n + 2 * p - dil * (k - 1) - 1) // stride + 1

    Xo, Yo, Zo = osz(X, KX), osz(Y, KY), osz(Z, KZ)
    out = torch.empty((B, Xo, Yo, Zo, Cout), dtype=torch.float32, device=x_cl.device)
    if gn_stats is not None:
        _chk(gn_stats, "gn_stats", torch.float64)
    ws = _conv_workspace(x_cl.device)
    check(lib().occ_conv_bf16x3(_ptr(x_cl), _ptr(w2), _ptr(out), B, X, Y, Z, Cin, Cout, KX, KY, KZ, stride, dil,
                                _ptr(bias), _ptr(residual), act, int(split_out), _ptr(gn_stats), cpg, _ptr(ws),
                                ws.numel(), _stream(x_cl)), "occ_conv_bf16x3")
    LAUNCH_COUNT[0] += 1
    return out


def conv_taps(x_cl, w2, taps, gn_stats=None, cpg=0):
    """Stride-1 convolution over an explicit tap table (several dilated branches in one launch): x_cl (B,X,Y,Z,Cin) S32,
    taps = [(dx, dy, dz), ...], w2 (Cout, len(taps)*Cin) S32 tap-major -> raw (B,X,Y,Z,Cout) fp32."""
    _chk(x_cl, "x_cl"), _chk(w2, "w2")
    B, X, Y, Z, Cin = x_cl.shape
    Cout = w2.shape[0]
    assert w2.shape[1] == len(taps) * Cin
    out = torch.empty((B, X, Y, Z, Cout), dtype=torch.float32, device=x_cl.device)
    flat = [int(v) for t in taps for v in t]
    arr = (ctypes.c_int * len(flat))(*flat)
    if gn_stats is not None:
        _chk(gn_stats, "gn_stats", torch.float64)
    check(lib().occ_conv_taps_bf16x3(_ptr(x_cl), _ptr(w2), _ptr(out), B, X, Y, Z, Cin, Cout, len(taps), arr, _ptr(gn_stats),
                                     cpg, _stream(x_cl)), "occ_conv_taps_bf16x3")
    LAUNCH_COUNT[0] += 1
    return out


# ----------------------------------------------------------------------------------------------- encoder glue
_WL = {}


def window_layout_buffer(B, X, Y, Z, C, shift, device):
    """Zero-initialised window-layout token buffer (occ_window_layout_rows, C), cached per (grid, shift, device, stream):
    the producer overwrites exactly the rows of real tokens on every call, the pad rows stay zero for ever."""
    key = (B, X, Y, Z, C, int(bool(shift)), str(device), torch.cuda.current_stream(device).cuda_stream)
    buf = _WL.get(key)
    if buf is None:
        rows = lib().occ_window_layout_rows(B, X, Y, Z)
        assert rows > 0
        buf = _WL[key] = torch.zeros((rows, C), dtype=torch.float32, device=device)
    return buf


def window_layout_index(B, X, Y, Z, shift):
    """Token row (voxel tokens, then BEV tokens) -> row of the window-layout buffer, as an int64 tensor (host arithmetic,
    independent of the kernels: tests and callers that hold token-ordered rows)."""
    WS = 7
    nWx, nWy = (X + WS - 1) // WS, (Y + WS - 1) // WS
    Xp, Yp = nWx * WS, nWy * WS
    x = torch.arange(X).view(X, 1)
    y = torch.arange(Y).view(1, Y)
    xs, ys = ((x - 3) % Xp, (y - 3) % Yp) if shift else (x, y)
    win_xy = (xs // WS) * nWy + (ys // WS)              # (X, Y)
    t = (xs % WS) * WS + (ys % WS)                      # (X, Y)
    b = torch.arange(B).view(B, 1, 1, 1)
    z = torch.arange(Z).view(1, 1, 1, Z)
    img = b * Z + z                                     # voxel slice image index
    nimg = B * (Z + 1)                                  # window index = (wx * nWy + wy) * n_images + image
    vox = (win_xy.view(1, X, Y, 1) * nimg + img) * 64 + t.view(1, X, Y, 1)                # (B, X, Y, Z)
    bev = (win_xy.view(1, X, Y) * nimg + (B * Z + torch.arange(B).view(B, 1, 1))) * 64 + t.view(1, X, Y)
    return torch.cat([vox.reshape(-1), bev.reshape(-1)]).long()


def to_window_layout(tokn_s, B, X, Y, Z, shift):
    """token-ordered (rows, C) rows -> window-layout buffer (fresh tensor)"""
    rows = lib().occ_window_layout_rows(B, X, Y, Z)
    out = torch.zeros((rows, tokn_s.shape[1]), dtype=tokn_s.dtype, device=tokn_s.device)
    out[window_layout_index(B, X, Y, Z, shift).to(tokn_s.device)] = tokn_s
    return out


def gn_relu_zmean_ln(y, stats, gn_w, gn_b, ln_w, ln_b, B, XY, Z, C, groups, X=0, win_shift=None):
    """-> tok (fp32, token order), tokn (S32): token order, or -- win_shift = 0 / 1 -- the window layout of that partition
    (operand of swin_qkv_attention)."""
    rows = B * XY * (Z + 1)
    tok = torch.empty((rows, C), dtype=torch.float32, device=y.device)
    if win_shift is None:
        tokn = torch.empty((rows, C), dtype=torch.float32, device=y.device)
        ws = -1
    else:
        assert X > 0 and XY % X == 0
        tokn = window_layout_buffer(B, X, XY // X, Z, C, win_shift, y.device)
        ws = int(bool(win_shift))
    check(lib().occ_gn_relu_zmean_ln(_ptr(y), _ptr(stats), _ptr(gn_w), _ptr(gn_b), _ptr(ln_w), _ptr(ln_b), _ptr(tok),
                                     _ptr(tokn), B, XY, Z, C, groups, X, ws, _stream()), "occ_gn_relu_zmean_ln")
    LAUNCH_COUNT[0] += 1
    return tok, tokn


def layernorm(x, w, b, split_out=False):
    _chk(x, "x")
    rows, C = x.shape
    out = torch.empty_like(x)
    check(lib().occ_layernorm(_ptr(x), _ptr(w), _ptr(b), _ptr(out), rows, C, int(split_out), _stream(x)), "occ_layernorm")
    LAUNCH_COUNT[0] += 1
    return out


def swin_proj_ffn(att, tok, wp, bp, ln_w, ln_b, w1, b1, w2, b2):
    """Fused Swin block tail (C == 128): tok + proj(att) -> LN2 -> FFN(GELU) + residual, one kernel."""
    _chk(att, "att"), _chk(tok, "tok")
    M, C = att.shape
    assert tok.shape == (M, C) and wp.shape == (C, C) and w1.shape == (C, C) and w2.shape == (C, C)
    out = torch.empty_like(att)
    check(lib().occ_swin_proj_ffn(_ptr(att), _ptr(tok), _ptr(wp), _ptr(bp), _ptr(ln_w), _ptr(ln_b), _ptr(w1), _ptr(b1),
                                  _ptr(w2), _ptr(b2), _ptr(out), M, C, _stream()), "occ_swin_proj_ffn")
    LAUNCH_COUNT[0] += 1
    return out


def gn_apply(x, stats, w, b, rows_per_batch, groups, residual=None, relu=True, want_f32=True, split_into=None,
             out_off=0, want_split=False, out_f32=None):
    """GroupNorm apply (+ReLU, +residual) of a raw conv output x (rows, C).  Returns (out_f32 | None, out_s32 | None):
    the fp32 result when want_f32, the S32 result written into split_into[:, out_off:out_off+C] (or a fresh (rows, C)
    tensor when want_split)."""
    _chk(x, "x")
    rows, C = x.shape
    out = out_f32 if out_f32 is not None else (torch.empty_like(x) if want_f32 else None)
    if out is not None:
        _chk(out, "out_f32")
        assert out.shape == x.shape
    if split_into is None and want_split:
        split_into = torch.empty_like(x)
    ldo = split_into.shape[1] if split_into is not None else C
    check(lib().occ_gn_apply(_ptr(x), _ptr(stats), _ptr(w), _ptr(b), _ptr(residual), _ptr(out), _ptr(split_into), rows,
                             rows_per_batch, C, groups, ldo, out_off, int(relu), _stream(x)), "occ_gn_apply")
    LAUNCH_COUNT[0] += 1
    return out, split_into


def aspp_gap_branch(x, wconv, gw, gb, cat, B, rows_per_batch, groups, out_off):
    ch = x.shape[1]
    sums = torch.empty((B, ch * 3 // 2 + 2), dtype=torch.float64, device=x.device)  # B*ch doubles + B*ch floats
    check(lib().occ_aspp_gap_branch(_ptr(x), _ptr(sums), _ptr(wconv), _ptr(gw), _ptr(gb), _ptr(cat), B, rows_per_batch,
                                    ch, groups, cat.shape[1], out_off, _stream()), "occ_aspp_gap_branch")
    LAUNCH_COUNT[0] += 4
    return cat


def dualpath_fuse(x, bev, cw, cbias, identity, B, XY, Z, C, identity_split=False, id_stats=None, id_w=None, id_b=None,
                  groups=0, want_f32=True, want_split=True):
    """-> (out fp32 | None, out S32 | None), each (B*XY*Z, C)."""
    out = torch.empty((B * XY * Z, C), dtype=torch.float32, device=x.device) if want_f32 else None
    out_s = torch.empty((B * XY * Z, C), dtype=torch.float32, device=x.device) if want_split else None
    check(lib().occ_dualpath_fuse(_ptr(x), _ptr(bev), _ptr(cw), float(cbias), _ptr(identity), int(identity_split),
                                  _ptr(id_stats), _ptr(id_w), _ptr(id_b), groups, _ptr(out), _ptr(out_s), B, XY, Z, C,
                                  _stream(x)), "occ_dualpath_fuse")
    LAUNCH_COUNT[0] += 1
    return out, out_s


def qkv_head_major_perm(C, heads, device=None):
    """row permutation of WindowMSA.qkv (weight / bias): new row h*96 + which*32 + d <- reference row which*C + h*32 + d."""
    hd = C // heads
    idx = torch.arange(3 * C, device=device).view(3, heads, hd).permute(1, 0, 2).reshape(-1)
    return idx


def window_attention(qkv, qkv_bias, bias_pad, B, X, Y, Z, C, heads, shift, head_major=False):
    _chk(qkv, "qkv")
    rows = B * X * Y * (Z + 1)
    assert qkv.shape == (rows, 3 * C)
    out = torch.empty((rows, C), dtype=torch.float32, device=qkv.device)
    check(lib().occ_window_attention(_ptr(qkv), _ptr(qkv_bias), _ptr(bias_pad), _ptr(out), B, X, Y, Z, C, heads,
                                     int(shift), int(head_major), _stream()), "occ_window_attention")
    LAUNCH_COUNT[0] += 1
    return out


def swin_qkv_attention(tokn_wl, w_qkv, b_qkv, bias_pad, B, X, Y, Z, C, heads, shift):
    """QKV projection + (shifted) window attention in one kernel (C == 128): tokn_wl (window_layout rows, C) S32 in the
    window layout of ``shift`` (gn_relu_zmean_ln(win_shift=shift) / to_window_layout), w_qkv (3C, C) S32 and b_qkv (3C,)
    fp32 with head-major rows -> attention output (B*X*Y*(Z+1), C) S32 in token order."""
    _chk(tokn_wl, "tokn_wl"), _chk(w_qkv, "w_qkv"), _chk(b_qkv, "b_qkv")
    rows = B * X * Y * (Z + 1)
    assert tokn_wl.shape == (lib().occ_window_layout_rows(B, X, Y, Z), C) and w_qkv.shape == (3 * C, C)
    out = torch.empty((rows, C), dtype=torch.float32, device=tokn_wl.device)
    check(lib().occ_swin_qkv_attention(_ptr(tokn_wl), _ptr(w_qkv), _ptr(b_qkv), _ptr(bias_pad), _ptr(out), B, X, Y, Z, C, heads,
                                       int(shift), _stream(tokn_wl)), "occ_swin_qkv_attention")
    LAUNCH_COUNT[0] += 1
    return out


# ----------------------------------------------------------------------------------------------- decoder head
def sine_pos3d(X, Y, Z, num_feats, device, temperature=10000.0, scale=6.283185307179586, eps=1e-6, offset=0.0):
    out = torch.empty((X * Y * Z, 3 * num_feats), dtype=torch.float32, device=device)
    check(lib().occ_sine_pos3d(_ptr(out), X, Y, Z, num_feats, temperature, scale, eps, offset, _stream()), "occ_sine_pos3d")
    LAUNCH_COUNT[0] += 1
    return out


def head_prep(x, channel_last, level_embed=None, pos=None):
    """x: (B,S,C) channel-last memory or (B,C,S) reference layout -> mem (B,S,C) [, kpos (B,S,C)] in S32."""
    _chk(x, "x")
    if channel_last:
        B, S, C = x.shape
    else:
        B, C, S = x.shape
    mem = torch.empty((B, S, C), dtype=torch.float32, device=x.device)
    kpos = torch.empty_like(mem) if pos is not None else None
    check(lib().occ_head_prep(_ptr(x), int(channel_last), _ptr(level_embed), _ptr(pos), _ptr(mem), _ptr(kpos), B, S, C,
                              _stream()), "occ_head_prep")
    LAUNCH_COUNT[0] += 1
    return mem, kpos


def query_head(query_in, W, NC, norm2=None, next_q=None, ffn_part=None):
    """forward_head query side.  norm2 = (w, b): query_in is the FFN accumulator, LN(norms.2) gives the layer output
    (returned as `query`); next_q = (query_pos, Q, wqT, bq, scale): also project the next layer's cross-attn queries."""
    rows, E = query_in.shape
    cls = torch.empty((rows, NC), dtype=torch.float32, device=query_in.device)
    membed = torch.empty((rows, E), dtype=torch.float32, device=query_in.device)
    query = torch.empty_like(query_in) if norm2 is not None else query_in
    qh = torch.empty_like(query_in) if next_q is not None else None
    n2w, n2b = norm2 if norm2 is not None else (None, None)
    qpos, Q, wqT, bq, scale = next_q if next_q is not None else (None, 0, None, None, 0.0)
    check(lib().occ_query_head(_ptr(query_in), _ptr(n2w), _ptr(n2b), _ptr(query) if norm2 is not None else None,
                               _ptr(W["pn_w"]), _ptr(W["pn_b"]), _ptr(W["clsT"]), _ptr(W["cls_b"]), NC, _ptr(W["m0T"]),
                               _ptr(W["m0b"]), _ptr(W["m1T"]), _ptr(W["m1b"]), _ptr(W["m2T"]), _ptr(W["m2b"]), _ptr(cls),
                               _ptr(membed), _ptr(qpos), Q, _ptr(wqT), _ptr(bq), scale, _ptr(qh), _ptr(ffn_part),
                               ffn_part.shape[0] if ffn_part is not None else 0, rows, E, _stream()),
          "occ_query_head")
    LAUNCH_COUNT[0] += 1
    return cls, membed, query, qh


def decode_ordered(t):
    """order-preserving int32 -> float32 (inverse of the kernels' encoding; used by tests / debugging)."""
    i = t.to(torch.int32)
    return torch.where(i >= 0, i, i ^ 0x7FFFFFFF).view(torch.float32)


def mask_pool(mask, B, grid, out_grid, Q):
    """General adaptive max pool of the query-last mask logits -> (pooled ordered-int32 (B,So,Q), flag int32 (B*Q,))."""
    X, Y, Z = grid
    Xo, Yo, Zo = out_grid
    pooled = torch.empty((B, Xo * Yo * Zo, Q), dtype=torch.int32, device=mask.device)
    flag = torch.empty((B * Q,), dtype=torch.int32, device=mask.device)
    check(lib().occ_mask_pool(_ptr(mask), _ptr(pooled), _ptr(flag), B, X, Y, Z, Xo, Yo, Zo, Q, _stream()), "occ_mask_pool")
    LAUNCH_COUNT[0] += 2
    return pooled, flag


def pool_fusable(grid, out_grid):
    """occ_mask_gemm_pool handles windows that are powers of two >= 2 dividing the grid."""
    for n, o in zip(grid, out_grid):
        if n % o:
            return False
        w = n // o
        if w < 2 or (w & (w - 1)):
            return False
    return True


def mask_gemm_pool(mf_r, membed, B, grid, out_grid, Q, want_mask):
    """mask logits (optional) + pooled attention-mask logits in one tensor-core kernel per sample."""
    X, Y, Z = grid
    Xo, Yo, Zo = out_grid
    E = mf_r.shape[-1]
    V = X * Y * Z
    mask = torch.empty((B, V, Q), dtype=torch.float32, device=mf_r.device) if want_mask else None
    pooled = torch.empty((B, Xo * Yo * Zo, Q), dtype=torch.int32, device=mf_r.device)
    flag = torch.empty((B * Q,), dtype=torch.int32, device=mf_r.device)
    check(lib().occ_mask_gemm_pool(_ptr(mf_r), _ptr(membed), _ptr(mask), _ptr(pooled), _ptr(flag), B, X, Y, Z, E, Q, Xo, Yo,
                                   Zo, _stream()), "occ_mask_gemm_pool")
    LAUNCH_COUNT[0] += 2 + B
    return mask, pooled, flag


def cross_attn_tc(qh, Kp, Vp, ld, koff, voff, pooled, flag, B, S, Q, E, H):
    """Masked cross attention on the tensor cores -> (partials (B,H,npart,Q,34), npart)."""
    npart = lib().occ_cross_attn_tc_partials(S)
    part = torch.empty((B, H, npart, Q, 34), dtype=torch.float32, device=qh.device)
    bits = torch.empty((B, 4 * ((S + 127) // 128), Q), dtype=torch.int32, device=qh.device)
    check(lib().occ_mask_bits(_ptr(pooled), _ptr(bits), B, S, Q, _stream()), "occ_mask_bits")
    check(lib().occ_cross_attn_tc(_ptr(qh), _ptr(Kp), _ptr(Vp), ld, koff, voff, _ptr(bits), _ptr(flag), _ptr(part), B, S, Q,
                                  E, H, _stream()), "occ_cross_attn_tc")
    LAUNCH_COUNT[0] += 2
    return part, npart


def cross_merge(part, nchunk, H, query, query_pos, Q, L, scale):
    rows, E = query.shape
    q1 = torch.empty_like(query)
    sa = torch.empty((rows, 3 * E), dtype=torch.float32, device=query.device)
    check(lib().occ_cross_merge(_ptr(part), nchunk, H, _ptr(query), _ptr(query_pos), Q, _ptr(L["ca_woT"]), _ptr(L["ca_bo"]),
                                _ptr(L["n0w"]), _ptr(L["n0b"]), _ptr(L["sa_inT"]), _ptr(L["sa_inb"]), scale, _ptr(q1),
                                _ptr(sa), rows, E, _stream()), "occ_cross_merge")
    LAUNCH_COUNT[0] += 1
    return q1, sa


def self_attn_ffn(sa, q1, Q, L, H):
    """-> (ybuf = x1 + b2, ffn_part (F/E, rows, E)): the pre-norms.2 accumulator of the decoder layer and the FFN
    column-block partials that occ_query_head adds to it in a fixed order."""
    rows, E = q1.shape
    x1 = torch.empty_like(q1)
    ybuf = torch.empty_like(q1)
    F = L["f1b"].numel()
    part = torch.empty((F // E, rows, E), dtype=torch.float32, device=q1.device)
    check(lib().occ_self_attn_ffn(_ptr(sa), _ptr(q1), Q, _ptr(L["sa_woT"]), _ptr(L["sa_bo"]), _ptr(L["n1w"]), _ptr(L["n1b"]),
                                  _ptr(L["f1T"]), _ptr(L["f1b"]), _ptr(L["f2T"]), _ptr(L["f2b"]), F, _ptr(x1), _ptr(ybuf),
                                  _ptr(part), rows, E, H, _stream()), "occ_self_attn_ffn")
    LAUNCH_COUNT[0] += 2
    return ybuf, part


def classmix(mask, cls, B, grid, out_grid, Q, NC, with_labels=False):
    X, Y, Z = grid
    Xo, Yo, Zo = out_grid
    out = torch.empty((B, NC - 1, Xo, Yo, Zo), dtype=torch.float32, device=mask.device)
    labels = torch.empty((B, Xo, Yo, Zo), dtype=torch.uint8, device=mask.device) if with_labels else None
    check(lib().occ_classmix(_ptr(mask), _ptr(cls), _ptr(out), _ptr(labels), B, X, Y, Z, Xo, Yo, Zo, Q, NC, _stream()),
          "occ_classmix")
    LAUNCH_COUNT[0] += 1
    return (out, labels) if with_labels else out


def transpose_sq(mask, B, S, Q):
    out = torch.empty((B, Q, S), dtype=torch.float32, device=mask.device)
    check(lib().occ_transpose_sq(_ptr(mask), _ptr(out), B, S, Q, _stream()), "occ_transpose_sq")
    LAUNCH_COUNT[0] += 1
    return out


def lidarseg_points(vox, pts, pc_range, border=True):
    """vox (K,X,Y,Z) contiguous class volume of one sample; pts (n, >=3) -> (n, K) softmaxed point scores."""
    _chk(vox, "vox")
    K, X, Y, Z = vox.shape
    pts = pts.float().contiguous()
    n = pts.shape[0]
    out = torch.empty((n, K), dtype=torch.float32, device=vox.device)
    r = [float(v) for v in pc_range]
    check(lib().occ_lidarseg_points(_ptr(vox), _ptr(pts) if n else None, pts.shape[1] if n else 3, n, r[0], r[1], r[2],
                                    r[3], r[4], r[5], X, Y, Z, K, int(border), _ptr(out), _stream()), "occ_lidarseg_points")
    LAUNCH_COUNT[0] += 1
    return out


# ----------------------------------------------------------------------------------------------- neck (pixel decoder)
def _grids_arr(grids):
    flat = [int(v) for g in grids for v in g]
    return (ctypes.c_int * len(flat))(*flat)


def epilogue_stats_plan(C, groups):
    """GroupNorm statistics can ride in the conv / GEMM epilogue when the group size is a power of two; otherwise (192
    channels / 32 groups: 6 = 2 * 3) the epilogue gathers them at the largest power-of-two sub-group and stats_regroup sums
    `factor` of those.  -> (cpg_epilogue, factor) or None when the epilogue cannot hold the sub-groups (> 96)."""
    cpg = C // groups
    sub = cpg & -cpg  # largest power-of-two divisor
    if sub > 32:
        sub = 32
    factor = cpg // sub
    return (sub, factor) if C // sub <= 96 else None


def stats_regroup(sub_stats, groups, factor):
    B = sub_stats.shape[0]
    if factor == 1:
        return sub_stats
    out = torch.empty((B, groups, 2), dtype=torch.float64, device=sub_stats.device)
    check(lib().occ_stats_regroup(_ptr(sub_stats), _ptr(out), B, groups, factor, _stream(sub_stats)), "occ_stats_regroup")
    LAUNCH_COUNT[0] += 1
    return out


def gn_stats(x, B, rows_per_batch, C, groups):
    """fp64 (sum, sumsq) per (sample, group) of x (B*rows_per_batch, C) -> (B, groups, 2); for the group sizes the conv
    epilogue does not accumulate (C / groups not a power of two)."""
    _chk(x, "x")
    stats = torch.zeros((B, groups, 2), dtype=torch.float64, device=x.device)
    check(lib().occ_gn_stats(_ptr(x), _ptr(stats), B, rows_per_batch, C, C // groups, _stream(x)), "occ_gn_stats")
    LAUNCH_COUNT[0] += 2
    return stats


def neck_token_prep(x, grids, B, ln=None, pos=None, want_f32=True, want_s32=True, want_pos=False):
    """[LayerNorm] of level-major token rows x (B*Nq, C) -> (x fp32 | None, x S32 | None, (x + pos) S32 | None)."""
    _chk(x, "x")
    rows, C = x.shape
    f32 = torch.empty_like(x) if want_f32 else None
    s32 = torch.empty_like(x) if want_s32 else None
    ps = torch.empty_like(x) if want_pos else None
    lw, lb = ln if ln is not None else (None, None)
    check(lib().occ_neck_token_prep(_ptr(x), _ptr(lw), _ptr(lb), _ptr(pos), _ptr(f32), _ptr(s32), _ptr(ps), len(grids), B,
                                    _grids_arr(grids), C, _stream(x)), "occ_neck_token_prep")
    LAUNCH_COUNT[0] += 1
    return f32, s32, ps


def ms_deform_attn(value, ow, grids, strides, B, E, H, P):
    """3-D multi-scale deformable attention core on level-major rows -> (rows, E) in S32.  value (rows, H * head_ld): head h
    in columns [h*head_ld, h*head_ld + E/H) -- head_ld = E/H, or E/H rounded up to 32 (``pad_head_rows``)."""
    _chk(value, "value"), _chk(ow, "ow")
    assert value.shape[1] % H == 0 and value.shape[1] // H >= E // H
    out = torch.empty(value.shape[0], E, dtype=torch.float32, device=value.device)
    st = (ctypes.c_float * len(strides))(*[float(v) for v in strides])
    check(lib().occ_ms_deform_attn(_ptr(value), value.shape[1], value.shape[1] // H, _ptr(ow), _ptr(out), len(grids), B,
                                   _grids_arr(grids), st, E, H, P, _stream(value)), "occ_ms_deform_attn")
    LAUNCH_COUNT[0] += 1
    return out


def pad_head_rows(w, b, H):
    """value_proj (E, K) weight / (E,) bias -> (H * hp, K) / (H * hp,) with hp = E/H rounded up to 32 and zero rows in the
    pad: the projection then writes every head slice into its own 128-byte line (see occ_ms_deform_attn)."""
    E = w.shape[0]
    hd = E // H
    hp = (hd + 31) // 32 * 32
    if hp == hd:
        return w.detach().float().contiguous(), b.detach().float().contiguous()
    wp = torch.zeros(H, hp, w.shape[1], dtype=torch.float32, device=w.device)
    wp[:, :hd] = w.detach().float().view(H, hd, -1)
    bp = torch.zeros(H, hp, dtype=torch.float32, device=w.device)
    bp[:, :hd] = b.detach().float().view(H, hd)
    return wp.view(H * hp, -1).contiguous(), bp.view(-1).contiguous()


def gn_upsample_add(cur, stats, gw, gb, groups, coarse):
    """cur (B,X,Y,Z,C) raw conv output (+ stats), coarse (B,Xc,Yc,Zc,C) fp32 -> GN(cur) + trilinear_up(coarse) in S32."""
    _chk(cur, "cur"), _chk(coarse, "coarse")
    B, X, Y, Z, C = cur.shape
    _, Xc, Yc, Zc, _ = coarse.shape
    out = torch.empty_like(cur)
    check(lib().occ_gn_upsample_add(_ptr(cur), _ptr(stats), _ptr(gw), _ptr(gb), groups, _ptr(coarse), _ptr(out), B, X, Y, Z,
                                    Xc, Yc, Zc, C, _stream(cur)), "occ_gn_upsample_add")
    LAUNCH_COUNT[0] += 1
    return out


# ----------------------------------------------------------------------------------------------- evaluation tail
def ssc_counts(pred, target, K, ignore=255):
    """uint8 label volumes (any shape) -> int64 (3 + 3K,): completion tp/fp/fn, semantic tp/fp/fn per class."""
    p = _chk(pred.reshape(-1).to(torch.uint8).contiguous(), "pred", torch.uint8)
    t = _chk(target.reshape(-1).to(torch.uint8).contiguous(), "target", torch.uint8)
    assert p.numel() == t.numel()
    ws = torch.empty(K * K, dtype=torch.int64, device=p.device)
    out = torch.empty(3 + 3 * K, dtype=torch.int64, device=p.device)
    check(lib().occ_ssc_counts(_ptr(p), _ptr(t), p.numel(), K, ignore, _ptr(ws), _ptr(out), _stream(p)), "occ_ssc_counts")
    LAUNCH_COUNT[0] += 2
    return out


def lidarseg_hist(scores, labels, K, hist=None):
    _chk(scores, "scores")
    labels = _chk(labels.long().contiguous(), "labels", torch.int64)
    if hist is None:
        hist = torch.zeros((K - 1, K - 1), dtype=torch.int64, device=scores.device)
    check(lib().occ_lidarseg_hist(_ptr(scores), _ptr(labels), scores.shape[0], K, _ptr(hist), _stream(scores)),
          "occ_lidarseg_hist")
    LAUNCH_COUNT[0] += 1
    return hist
