"""Per-kernel timings at BASELINE sizes (CUDA events, L2 flushed between iterations).  Development aid;
bench.py is the contract benchmark."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from occformer_b200 import ops, synth
from occformer_b200.view_transformer import ViewTransformerLiftSplatShootVoxel

dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    res = {}
    which = sys.argv[1:] or ["pool", "fill", "gemm", "conv", "attn", "block"]
    X, Y, Z, C = 200, 200, 16, 128
    if "pool" in which:
        from oracle import port  # geometry only (inputs), not on the measured path
        gc = synth.grid_config("nusc_200")
        vt = ViewTransformerLiftSplatShootVoxel(grid_config=gc, data_config={"input_size": (256, 704)}, numC_Trans=C).to(dev)
        cams = {k: v.to(dev) for k, v in synth.nusc_cameras(1, 6).items()}
        geom = vt.get_geometry(**cams).contiguous()
        dd, feat = synth.lift_inputs(1, 6, 112, 16, 44, C)
        dd, feat = dd.to(dev), feat.to(dev)
        prob, feat_cl = ops.lift_prologue(dd, feat)
        dx, bx, nx = vt._host_params()
        t = timeit(lambda: ops.lift_splat(prob, feat_cl, geom, 1, 6, dx, bx, nx, vt.grid_size()))
        P = geom.numel() // 3
        alg = P * 4 + 6 * 16 * 44 * C * 4 + P * 12 + X * Y * Z * C * 4
        res["lift_splat_ms"] = t
        res["lift_splat_GBps_fused_formula"] = alg / t / 1e6
        t2 = timeit(lambda: ops.lift_prologue(dd, feat))
        res["lift_prologue_ms"] = t2
    if "fill" in which:
        big = torch.empty(1, X, Y, Z, C, device=dev)
        t = timeit(lambda: big.zero_())
        res["torch_zero_328MB_ms"] = t
        res["torch_zero_GBps"] = big.numel() * 4 / t / 1e6
        src = torch.randn(1, X, Y, Z, C, device=dev)
        t = timeit(lambda: big.copy_(src))
        res["torch_copy_328MB_ms"] = t
        res["torch_copy_GBps(read+write)"] = 2 * big.numel() * 4 / t / 1e6
    if "gemm" in which:
        M = X * Y * (Z + 1)
        a = torch.randn(M, C, device=dev)
        for N in (384, 128):
            w = torch.randn(N, C, device=dev)
            b = torch.randn(N, device=dev)
            t = timeit(lambda: ops.gemm(a, w, bias=b))
            res[f"gemm_M{M}_N{N}_K{C}_ms"] = t
            res[f"gemm_M{M}_N{N}_K{C}_TFLOPs"] = 2.0 * M * N * C / t / 1e9
    if "conv" in which:
        x = torch.randn(1, X, Y, Z, C, device=dev)
        w2, ks = ops.repack_conv_weight(torch.randn(C, C, 3, 3, 3, device=dev))
        stats = torch.zeros(1, 32, 2, dtype=torch.float64, device=dev)
        t = timeit(lambda: ops.conv(x, w2, ks, gn_stats=stats, cpg=4))
        res["conv3d_200x200x16_c128_ms"] = t
        res["conv3d_TFLOPs"] = 2.0 * 27 * C * C * X * Y * Z / t / 1e9
        xl = torch.randn(1, X, Y, Z, C, device=dev).permute(0, 4, 1, 2, 3).contiguous(memory_format=torch.channels_last_3d)
        wt = torch.randn(C, C, 3, 3, 3, device=dev).contiguous(memory_format=torch.channels_last_3d)
        t = timeit(lambda: torch.nn.functional.conv3d(xl, wt, padding=1))
        res["cudnn_conv3d_ms(for context)"] = t
    if "attn" in which:
        M = X * Y * (Z + 1)
        qkv = torch.randn(M, 3 * C, device=dev)
        qb = torch.randn(3 * C, device=dev)
        bd = torch.randn(4, 2404, device=dev)
        t = timeit(lambda: ops.window_attention(qkv, qb, bd, 1, X, Y, Z, C, 4, True))
        res["window_attn_tc_ms"] = t
        res["window_attn_tc_head_major_ms"] = timeit(lambda: ops.window_attention(qkv, qb, bd, 1, X, Y, Z, C, 4, True, head_major=True))
        res["window_attn_GBps"] = (M * 3 * C * 4 + M * C * 4) / t / 1e6
    if "swin" in which:
        M = X * Y * (Z + 1)
        att = ops.round_tf32_(torch.randn(M, C, device=dev))
        tok = torch.randn(M, C, device=dev)
        ws = [ops.round_tf32_(torch.randn(C, C, device=dev) * C ** -0.5) for _ in range(3)]
        bs = [0.1 * torch.randn(C, device=dev) for _ in range(3)]
        lw, lb = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        t = timeit(lambda: ops.swin_proj_ffn(att, tok, ws[0], bs[0], lw, lb, ws[1], bs[1], ws[2], bs[2]))
        res["swin_proj_ffn_fused_ms"] = t
        res["swin_proj_ffn_GBps"] = 3 * M * C * 4 / t / 1e6
    if "block" in which:
        from occformer_b200.encoder import OccupancyEncoder
        enc = OccupancyEncoder(in_channels=128, num_stage=4, block_numbers=[2, 2, 2, 2], block_inplanes=[128, 256, 512, 1024],
                               block_strides=[1, 2, 2, 2], out_indices=(0, 1, 2, 3), norm_cfg=dict(type="GN", num_groups=32)).to(dev).eval()
        x = torch.randn(1, X, Y, Z, C, device=dev)
        t = timeit(lambda: enc.forward_cl(x), iters=3, warm=1)
        res["encoder_200x200x16_ms"] = t
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
