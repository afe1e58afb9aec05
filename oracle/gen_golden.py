"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REAL reference modules
(/root/reference imported verbatim under oracle/shim.py) on the deterministic synthetic inputs of
occformer_b200/synth.py with the deterministic weights of oracle/port.make_*_state.

    python -m oracle.gen_golden        # run in the build container (needs /root/reference)

The fixtures are what pins oracle/port.py (and, through it, the CUDA path) on the GPU box, where
/root/reference does not exist.  Inputs are NOT stored (they are regenerated from seeds); each
fixture stores the reference OUTPUT plus the recipe (shapes, seeds) needed to rebuild the input.
"""
import os
import sys

import numpy as np
import torch

from occformer_b200 import synth

from . import port, refmodels, shim

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

BLOCK_CASES = [
    # name, cin, c, stride, shift, grid, seed
    ("block_c128_s1_plain", 128, 128, 1, False, (15, 10, 4), 1),
    ("block_c256_s2_shift", 128, 256, 2, True, (15, 10, 4), 2),
    ("block_c128_s1_shift", 128, 128, 1, True, (9, 16, 2), 3),
]
HEAD_CASE = dict(E=96, Q=12, K=17, L=4, ffn=192, sizes=[(16, 12, 4), (8, 6, 2), (4, 3, 1), (2, 2, 1)],
                 occ_size=[32, 24, 8], pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], npts=50,
                 wseed=7, xseed=9, pseed=11)
POOL_CASE = dict(grid="pr1", input_size=(128, 128), B=2, N=1, C=32, seed=1)


def gen_neck():
    """MSDeformAttnPixelDecoder3D: reference outputs for validate_port.NECK_CASE, B = 1."""
    c, neck_inputs = port.NECK_CASE, port.neck_inputs
    sd = port.make_neck_state(c["in_channels"], c["E"], c["layers"], c["heads"], c["levels"], c["points"], c["ffn"],
                              seed=c["wseed"])
    neck = refmodels.build_neck(c["in_channels"], c["strides"], c["E"], c["layers"], c["heads"], c["levels"],
                                c["points"], c["ffn"], sd)
    feats = neck_inputs(c, B=1)
    with torch.no_grad():
        outs = neck([f.clone() for f in feats])
    np.savez(os.path.join(OUT, "neck_small.npz"), **{f"out{i}": o.numpy() for i, o in enumerate(outs)})
    print("neck_small", [tuple(o.shape) for o in outs])


DEPTHNET_CASE = dict(cin=32, mid=32, ctx=16, D=12, cam=27, B=1, N=2, fH=6, fW=9, wseed=31, xseed=33)


def depthnet_state(ref_module, seed):
    """The reference DepthNet's own state_dict with every tensor made non-trivial (BN running statistics, the
    zero-initialised DCN offset conv)."""
    g = torch.Generator().manual_seed(seed)
    sd = ref_module.state_dict()
    for k, v in sd.items():
        if "num_batches_tracked" in k:
            continue
        if "running_var" in k:
            v.copy_(0.5 + torch.rand(v.shape, generator=g))
        elif "running_mean" in k:
            v.copy_(0.2 * torch.randn(v.shape, generator=g))
        elif "conv_offset" in k:
            v.copy_(0.3 * torch.randn(v.shape, generator=g))
        else:
            v.copy_(torch.randn(v.shape, generator=g) * (0.1 if v.dim() == 1 else (v[0].numel()) ** -0.5))
            if v.dim() == 1 and ("bn" in k or k.endswith("1.weight")) and k.endswith("weight"):
                v.add_(1.0)
    return sd


def gen_depthnet():
    """DepthNet: reference class under the shim (mmdet BasicBlock / mmcv DCN restated in oracle/shim.py),
    its state_dict KEYS + SHAPES (the checkpoint contract) and its output on a seeded input."""
    import json
    c = DEPTHNET_CASE
    base = shim.load(refmodels.OCC + "image2bev.ViewTransformerLSSBEVDepth")
    torch.manual_seed(c["wseed"])
    ref = base.DepthNet(c["cin"], c["mid"], c["ctx"], c["D"], cam_channels=c["cam"]).eval()
    sd = depthnet_state(ref, c["wseed"])
    ref.load_state_dict(sd)
    g = torch.Generator().manual_seed(c["xseed"])
    x = torch.randn(c["B"] * c["N"], c["cin"], c["fH"], c["fW"], generator=g)
    mlp = torch.randn(c["B"], c["N"], c["cam"], generator=g)
    with torch.no_grad():
        y = ref(x, mlp)
    np.savez(os.path.join(OUT, "depthnet_small.npz"), out=y.numpy(), **{"w:" + k: v.numpy() for k, v in sd.items()})
    # the key / shape contract of the full-size module of the nuScenes config (numC_input = 512, numC_Trans = 128, D = 112)
    full = base.DepthNet(512, 512, 128, 112, cam_channels=27)
    with open(os.path.join(OUT, "depthnet_keys_nusc.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in full.state_dict().items()}, f, indent=0)
    print("depthnet_small", tuple(y.shape), len(sd), "keys")


def gen_model_cfg():
    """The ``model`` dict of the UNMODIFIED reference configs as plain data (the config files are pure-python assignments;
    ``_base_`` inheritance only adds dataset / runtime keys): the sections our registries must build from."""
    import json
    for name, rel in (("nusc_r50", "projects/configs/occformer_nusc/occformer_nusc_r50_256x704.py"),
                      ("kitti", "projects/configs/occformer_kitti/occformer_kitti.py")):
        ns = {}
        exec(compile(open(os.path.join(shim.REFERENCE_ROOT, rel)).read(), rel, "exec"), ns)
        m = ns["model"]
        keep = {k: m[k] for k in ("img_view_transformer", "img_bev_encoder_backbone", "img_bev_encoder_neck", "pts_bbox_head")}
        with open(os.path.join(OUT, f"model_cfg_{name}.json"), "w") as f:
            json.dump(keep, f, indent=1, default=lambda o: list(o))
        print("model_cfg", name, list(keep))


def main():
    assert shim.reference_available(), "needs /root/reference"
    shim.install()
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    if "neck" in sys.argv[1:]:  # python -m oracle.gen_golden neck : only the neck fixture
        gen_neck()
        return
    if "depthnet" in sys.argv[1:]:
        gen_depthnet()
        return
    if "cfg" in sys.argv[1:]:
        gen_model_cfg()
        return

    # ---- voxel pooling (reference ViewTransformerLiftSplatShootVoxel + its own QuickCumsum fallback)
    pc = POOL_CASE
    gc = synth.grid_config(pc["grid"])
    vt = refmodels.build_view_transformer(gc, pc["input_size"], numC_Trans=pc["C"])
    cams = synth.pr1_camera(B=pc["B"])
    geom = vt.get_geometry(cams["rots"], cams["trans"], cams["intrins"], cams["post_rots"],
                           cams["post_trans"], cams["bda"])
    D, fH, fW = vt.frustum.shape[:3]
    dd, feat = synth.lift_inputs(pc["B"], pc["N"], D, fH, fW, pc["C"], seed=pc["seed"])
    out, prob = refmodels.ref_lift_and_pool(vt, dd, feat, geom, pc["B"], pc["N"])
    dense = out.permute(0, 2, 3, 4, 1).contiguous()  # (B,X,Y,Z,C)
    nz = dense.abs().sum(-1) != 0
    np.savez(os.path.join(OUT, "voxel_pool_pr1.npz"), geom=geom.numpy(), nonzero_index=torch.nonzero(nz).numpy().astype(np.int32),
             nonzero_rows=dense[nz].numpy(), shape=np.array(dense.shape), depth_prob_sum=float(prob.double().sum()))
    print("voxel_pool_pr1: nonzero voxels", int(nz.sum()))

    # ---- encoder blocks
    for name, cin, c, stride, shift, grid, seed in BLOCK_CASES:
        g = torch.Generator().manual_seed(seed)
        sd = port.make_block_state(cin, c, stride, g)
        blk = refmodels.build_block(cin, c, stride, 1 if shift else 0, sd)
        x = synth.encoder_input(1, cin, *grid, seed=seed + 100)
        with torch.no_grad():
            y = blk(x.clone())
        np.savez(os.path.join(OUT, name + ".npz"), out=y.numpy(),
                 recipe=np.array([cin, c, stride, int(shift), *grid, seed]))
        print(name, tuple(y.shape))

    # ---- head
    hc = HEAD_CASE
    sd = port.make_head_state(hc["E"], hc["Q"], hc["K"], hc["L"], 3, ffn=hc["ffn"], seed=hc["wseed"])
    head = refmodels.build_head(hc["E"], hc["Q"], hc["K"], hc["L"], 3, hc["ffn"], sd)
    feats = synth.head_inputs(1, hc["E"], hc["sizes"], seed=hc["xseed"])
    metas = [dict(occ_size=hc["occ_size"], pc_range=hc["pc_range"])]
    pts = [synth.lidar_points(hc["npts"], hc["pc_range"], seed=hc["pseed"])]
    with torch.no_grad():
        cl, ml = head([f.clone() for f in feats], metas)
        res = head.simple_test([f.clone() for f in feats], metas, points=pts)
    np.savez(os.path.join(OUT, "head_nusc.npz"), cls=torch.stack(cl).numpy(), mask_last=ml[-1].numpy(),
             mask_first=ml[0].numpy(), output_voxels=res["output_voxels"][0].numpy(),
             output_points=res["output_points"].numpy())
    print("head_nusc done")
    gen_neck()
    gen_depthnet()
    gen_model_cfg()


if __name__ == "__main__":
    main()
