"""TEST INFRASTRUCTURE ONLY -- validates oracle/port.py against the REAL reference code
(imported verbatim under oracle/shim.py).  Runs only where /root/reference exists.

    python -m oracle.validate_port            # prints max abs / rel differences, asserts tight bounds
"""
import sys

import torch

from occformer_b200 import synth

from . import port, refmodels, shim


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def check_voxel_pooling(report):
    gc = synth.grid_config("pr1")
    vt = refmodels.build_view_transformer(gc, (128, 128), numC_Trans=32)
    cams = synth.pr1_camera(B=2)
    B, N = 2, 1
    geom_ref = vt.get_geometry(cams["rots"], cams["trans"], cams["intrins"], cams["post_rots"],
                               cams["post_trans"], cams["bda"])
    frustum = port.create_frustum((128, 128), 16, gc["dbound"])
    assert torch.equal(frustum, vt.frustum.data)
    geom = port.get_geometry(frustum, **cams)
    assert torch.equal(geom, geom_ref), "get_geometry differs"
    dx, bx, nx = port.gen_dx_bx(gc["xbound"], gc["ybound"], gc["zbound"])
    assert torch.equal(dx, vt.dx.data) and torch.equal(bx, vt.bx.data) and torch.equal(nx, vt.nx.data)
    D, fH, fW = frustum.shape[:3]
    dd, feat = synth.lift_inputs(B, N, D, fH, fW, 32, seed=1)
    (ref_out, ref_prob) = refmodels.ref_lift_and_pool(vt, dd, feat, geom_ref, B, N)
    vol, prob = port.lift(dd, feat, B, N)
    out, gf, kept = port.voxel_pooling(geom, vol, dx, bx, nx)
    assert torch.equal(prob, ref_prob)
    assert out.shape == ref_out.shape and out.stride() == ref_out.stride()
    r = rel(out, ref_out)
    report("voxel_pooling (pr1, B=2)", r, 2e-6, extra=f"kept {int(kept.sum())}/{kept.numel()}")


def check_block(report, cin, c, stride, shift, grid, seed):
    g = torch.Generator().manual_seed(seed)
    sd = port.make_block_state(cin, c, stride, g)
    blk = refmodels.build_block(cin, c, stride, 1 if shift else 0, sd)
    x = synth.encoder_input(1, cin, *grid, seed=seed + 100)
    with torch.no_grad():
        ref = blk(x.clone())
        out = port.dualpath_block(x, sd, "", stride, shift)
    report(f"dualpath_block cin={cin} c={c} s={stride} shift={shift} grid={grid}", rel(out, ref), 1e-5)


def check_encoder(report):
    sd = port.make_encoder_state(128, [128, 256], [2, 1], [1, 2], seed=3)
    enc = refmodels.build_encoder(128, [128, 256], [2, 1], [1, 2], (0, 1), sd)
    x = synth.encoder_input(1, 128, 15, 10, 4, seed=5)
    with torch.no_grad():
        ref = enc(x.clone())
        out = port.occupancy_encoder(x, sd, [2, 1], [1, 2], (0, 1))
    for i, (a, b) in enumerate(zip(out, ref)):
        report(f"occupancy_encoder out[{i}] {tuple(a.shape)}", rel(a, b), 1e-5)


def check_head(report, kitti=False):
    E, Q, K, L = 96, 12, 17, 4
    sd = port.make_head_state(E, Q, K, L, 3, ffn=2 * E, seed=7)
    head = refmodels.build_head(E, Q, K, L, 3, 2 * E, sd, kitti=kitti)
    sizes = [(16, 12, 4), (8, 6, 2), (4, 3, 1), (2, 2, 1)]
    feats = synth.head_inputs(1, E, sizes, seed=9)
    metas = [dict(occ_size=[32, 24, 8], pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0])]
    pts = [synth.lidar_points(50, metas[0]["pc_range"], seed=11)]
    with torch.no_grad():
        cl_ref, ml_ref = head([f.clone() for f in feats], metas)
        cl, ml = port.head_forward(feats, sd, E // 32, L, 3)
        for i in range(L + 1):
            report(f"head{'(kitti)' if kitti else ''} cls[{i}]", rel(cl[i], cl_ref[i]), 2e-5)
            report(f"head{'(kitti)' if kitti else ''} mask[{i}]", rel(ml[i], ml_ref[i]), 2e-5)
        if not kitti:
            res_ref = head.simple_test([f.clone() for f in feats], metas, points=pts)
            res = port.head_simple_test(feats, sd, E // 32, L, metas[0]["occ_size"], 3, points=pts,
                                        pc_range=metas[0]["pc_range"])
            report("head simple_test output_voxels", rel(res["output_voxels"][0], res_ref["output_voxels"][0]), 2e-5)
            report("head simple_test output_points", rel(res["output_points"], res_ref["output_points"]), 2e-5)


NECK_CASE, neck_inputs = port.NECK_CASE, port.neck_inputs


def check_neck(report):
    c = NECK_CASE
    sd = port.make_neck_state(c["in_channels"], c["E"], c["layers"], c["heads"], c["levels"], c["points"], c["ffn"],
                              seed=c["wseed"])
    neck = refmodels.build_neck(c["in_channels"], c["strides"], c["E"], c["layers"], c["heads"], c["levels"],
                                c["points"], c["ffn"], sd)
    feats = neck_inputs(c, B=2)
    with torch.no_grad():
        ref = neck([f.clone() for f in feats])
        out = port.ms_deform_pixel_decoder_3d(feats, sd, c["strides"], c["heads"], c["layers"], c["levels"], c["points"])
    assert len(ref) == len(out)
    for i, (a, b) in enumerate(zip(out, ref)):
        report(f"MSDeformAttnPixelDecoder3D out[{i}] {tuple(a.shape)}", rel(a, b), 2e-5)


def main():
    assert shim.reference_available(), "needs /root/reference"
    shim.install()
    torch.manual_seed(0)
    worst = []

    def report(name, r, tol, extra=""):
        ok = r <= tol
        worst.append(ok)
        print(f"{'OK ' if ok else 'BAD'} {name}: rel {r:.3e} (tol {tol:.0e}) {extra}")

    check_voxel_pooling(report)
    check_block(report, 128, 128, 1, False, (15, 10, 4), 1)
    check_block(report, 128, 256, 2, True, (15, 10, 4), 2)
    check_block(report, 128, 128, 1, True, (9, 16, 2), 3)
    check_encoder(report)
    check_head(report)
    check_head(report, kitti=True)
    check_neck(report)
    if not all(worst):
        sys.exit(1)
    print("port == reference on all checks")


if __name__ == "__main__":
    main()
