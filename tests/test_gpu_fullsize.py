"""Full-size parity of the benchmarked pipeline: ONE sample of the exact modules / weights / inputs bench.py times,
through the registered modules' ``forward`` (view transformer -> OccupancyEncoder -> [neck] -> head forward +
simple_test), against the CPU oracle (oracle/port.py, pinned to the reference) on the host.

  * nusc_200: BASELINE.json configs[2]  (6 cams 256x704 -> 200x200x16, occ 200x200x16)      = bench.py's workload
  * nusc_ref: the reference's own grid  (128x128x16 -> occ 256x256x32, occformer_nusc_r50_256x704.py:17-20,41-46)
  * kitti   : BASELINE.json configs[1]  (1 cam 384x1280, 4x4 P2 intrinsics, 4x4 bda, K = 20, Mask2FormerOccHead,
              128x128x16 -> occ 256x256x32, occformer_kitti.py)

Gate = SURVEY.md 8(d), both criteria, per output tensor (tests/util.py::assert_close); voxel bookkeeping bit exact;
bool attention-mask flips counted per decoder layer: layer 0 (kernel error only) gated, later layers (the discontinuous
threshold cascades) logged and gated through the tensors they feed.
"""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import port
from occformer_b200 import synth
from util import assert_close

pytestmark = pytest.mark.gpu

DOWNSAMPLE, C_TRANS = 16, 128
PLANES, NUMS, STRIDES = [128, 256, 512, 1024], [2, 2, 2, 2], [1, 2, 2, 2]
EMBED, QUERIES, DEC_LAYERS, HEADS = 192, 100, 9, 6
NECK = dict(strides=[2, 4, 8, 16], layers=6, heads=8, levels=3, points=4, ffn=4 * EMBED)

CASES = synth.WORKLOADS


class _PassThroughDepthNet(torch.nn.Module):
    """The view transformer's boundary is fed post-DepthNet maps (DepthNet needs mmcv DCN and is
    outside the replaced subsystems), so ``depth_net(x, mlp_input)`` hands x through on both arms."""

    def forward(self, x, mlp_input=None):
        return x


def _neck_available():
    try:
        from occformer_b200 import neck  # noqa: F401
        return True
    except ImportError:
        return False


def _oracle(case, x, cams, sd_e, sd_n, sd_h, head_feats):
    gc = synth.grid_config(case["grid"])
    N_CAMS, INPUT_SIZE = case["cams"], case["input_size"]
    frustum = port.create_frustum(INPUT_SIZE, DOWNSAMPLE, gc["dbound"])
    D = frustum.shape[0]
    geom = port.get_geometry(frustum, **cams)
    vol, prob = port.lift(x[:, :D], x[:, D:], 1, N_CAMS)
    dx, bx, nx = port.gen_dx_bx(gc["xbound"], gc["ybound"], gc["zbound"])
    grid, gf, kept = port.voxel_pooling(geom, vol, dx, bx, nx)
    del vol
    enc = port.occupancy_encoder(grid, sd_e, NUMS, STRIDES, (0, 1, 2, 3))
    if sd_n is not None:
        feats = port.ms_deform_pixel_decoder_3d(enc, sd_n, NECK["strides"], NECK["heads"], NECK["layers"], NECK["levels"],
                                                NECK["points"])
    else:
        feats = head_feats
    cls_l, mask_l = port.head_forward(feats, sd_h, HEADS, DEC_LAYERS, 3)
    res = port.head_simple_test(feats, sd_h, HEADS, DEC_LAYERS, case["occ"], 3)
    return dict(grid=grid, kept=int(kept.sum()), enc=enc, feats=feats, cls=cls_l, mask=mask_l,
                vox=res["output_voxels"][0], prob=prob)


@pytest.mark.parametrize("name", ["nusc_200", "nusc_ref", "kitti"])
def test_full_size_pipeline_vs_oracle(cuda, name):
    from occformer_b200 import BACKBONES, HEADS as HEAD_REG, NECKS
    from occformer_b200.head import head_cfg
    case = CASES[name]
    N_CAMS, INPUT_SIZE, CLASSES = case["cams"], case["input_size"], case["classes"]
    threads = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(threads)
    gc = synth.grid_config(case["grid"])
    X, Y, Z = (int(round((b[1] - b[0]) / b[2])) for b in (gc["xbound"], gc["ybound"], gc["zbound"]))
    fH, fW = INPUT_SIZE[0] // DOWNSAMPLE, INPUT_SIZE[1] // DOWNSAMPLE
    D = 112
    dd, feat = synth.lift_inputs(1, N_CAMS, D, fH, fW, C_TRANS, seed=0)
    x = torch.cat([dd, feat], dim=1).contiguous()
    cams = synth.workload_cameras(name, 1)
    sd_e = synth.make_encoder_state(C_TRANS, PLANES, NUMS, STRIDES, seed=0)
    sd_h = synth.make_head_state(EMBED, QUERIES, CLASSES, DEC_LAYERS, 3, seed=1)
    connected = _neck_available()
    sd_n = port.make_neck_state(PLANES, EMBED, NECK["layers"], NECK["heads"], NECK["levels"], NECK["points"], NECK["ffn"],
                                seed=2) if connected else None
    sizes = [(X, Y, Z), (X // 2, Y // 2, Z // 2), (X // 4, Y // 4, Z // 4), (X // 8, Y // 8, Z // 8)]
    head_feats = None if connected else synth.head_inputs(1, EMBED, sizes, seed=5)

    # ------------------------------------------------------------------ CUDA path through the registered modules
    vt = NECKS.build(dict(type="ViewTransformerLiftSplatShootVoxel", loss_depth_weight=1.0, grid_config=gc,
                          data_config={"input_size": INPUT_SIZE}, numC_input=D + C_TRANS, numC_Trans=C_TRANS,
                          depth_net=_PassThroughDepthNet())).to(cuda)
    enc = BACKBONES.build(dict(type="OccupancyEncoder", in_channels=C_TRANS, num_stage=4, block_numbers=NUMS,
                               block_inplanes=PLANES, block_strides=STRIDES, out_indices=(0, 1, 2, 3),
                               norm_cfg=dict(type="GN", num_groups=32, requires_grad=True), with_cp=True))
    enc.load_state_dict(sd_e, strict=True)
    enc = enc.to(cuda).eval()
    head = HEAD_REG.build(dict(type=case["head"], **head_cfg(EMBED, QUERIES, CLASSES, DEC_LAYERS, HEADS,
                                                                        case["pc"])))
    head.load_state_dict(sd_h, strict=True)
    head = head.to(cuda).eval()
    inp = [x.view(1, N_CAMS, D + C_TRANS, fH, fW).to(cuda)] + [cams[k].to(cuda) for k in
                                                              ("rots", "trans", "intrins", "post_rots", "post_trans", "bda")]
    inp.append(vt.get_mlp_input(*inp[1:7]) if hasattr(vt, "get_mlp_input") else None)
    voxel, depth_prob = vt(inp)
    enc_outs = enc(voxel)
    if connected:
        from occformer_b200.neck import neck_cfg
        neck = NECKS.build(dict(type="MSDeformAttnPixelDecoder3D", **neck_cfg(PLANES, NECK["strides"], EMBED, NECK["layers"],
                                                                              NECK["heads"], NECK["levels"], NECK["points"],
                                                                              NECK["ffn"])))
        neck.load_state_dict(sd_n, strict=True)
        neck = neck.to(cuda).eval()
        feats = neck(enc_outs)
    else:
        feats = [f.to(cuda) for f in head_feats]
    metas = [dict(occ_size=case["occ"], pc_range=case["pc"])]
    cls_l, mask_l = head(feats, metas)
    mask_cpu = [m.float().cpu() for m in mask_l]
    del mask_l
    res = head.simple_test(feats, metas)
    torch.cuda.synchronize()

    # ------------------------------------------------------------------ oracle on the host cores
    ref = _oracle(case, x, cams, sd_e, sd_n, sd_h, head_feats)

    tag = f"[{name}{' connected' if connected else ''}]"
    assert_close(depth_prob, ref["prob"], what=f"{tag} depth_prob")
    # voxel pooling: set of non-empty voxels identical, sums to fp32 summation-order accuracy
    g_cuda = voxel.float().cpu()
    assert torch.equal(g_cuda.abs().sum(1) != 0, ref["grid"].abs().sum(1) != 0), "non-empty voxel sets differ"
    # (1e-4: the summation order inside a voxel is unspecified -- the reference's argsort is unstable -- and depth_prob
    # itself carries the ~1e-6 of an fp32 exp; a few near-cancelling sums out of 1e8 exceed 1e-5 * rms)
    assert_close(g_cuda, ref["grid"], 1e-4, f"{tag} pooled voxel grid {tuple(ref['grid'].shape)}")
    for i, (o, r) in enumerate(zip(enc_outs, ref["enc"])):
        assert_close(o, r, what=f"{tag} encoder out[{i}] {tuple(r.shape)}")
    if connected:
        for i, (o, r) in enumerate(zip(feats, ref["feats"])):
            assert_close(o, r, what=f"{tag} neck out[{i}] {tuple(r.shape)}")
    # first and final predictions at the SURVEY gate; the mid-decoder prediction sits downstream of the discontinuous
    # attention-mask threshold (see below): logged, gated at 5e-3 (measured 1e-5 .. 4e-4 depending on which bits flip)
    assert_close(cls_l[0], ref["cls"][0], what=f"{tag} cls_pred[0]")
    assert_close(cls_l[DEC_LAYERS // 2], ref["cls"][DEC_LAYERS // 2], 5e-3, f"{tag} cls_pred[{DEC_LAYERS // 2}]")
    assert_close(cls_l[DEC_LAYERS], ref["cls"][DEC_LAYERS], what=f"{tag} cls_pred[{DEC_LAYERS}]")
    # bool attention masks of every decoder layer (mask2former_nusc_occ.py:463-466).  A flip needs |pooled logit| <= |error
    # of that logit| (the signs differ).  Layer 0 is predicted from the learned query embedding, so only kernel error can
    # flip its mask: gated at 3e-4 * max|logit| (measured 2e-5 .. 7e-5 = the logits' own rel_max).  From layer 1 on a
    # flipped mask bit changes the queries themselves and the next layer's logits move by more than kernel error -- the
    # threshold is discontinuous in the reference too (any fp32 reordering of the oracle cascades the same way; the run to
    # run order of the pooled sums is enough to change which bits flip) -- so those layers are counted and logged, and
    # gated through the tensors they feed: cls_pred[4], cls_pred[9], mask_pred[9], output_voxels, label agreement.
    nflip, worst, worst0 = 0, 0.0, 0.0
    for i in range(DEC_LAYERS):
        tgt = sizes[1:][::-1][i % 3]
        pa = F.adaptive_max_pool3d(mask_cpu[i], tgt)
        pb = F.adaptive_max_pool3d(ref["mask"][i], tgt)
        flips = (pa < 0) != (pb < 0)
        if flips.any():
            nflip += int(flips.sum())
            w = float(pb[flips].abs().max() / pb.abs().max())
            worst = max(worst, w)
            if i == 0:
                worst0 = w
    line = (f"[parity] {tag} attention-mask flips over {DEC_LAYERS} layers: {nflip}, largest |logit|/max|logit| among them "
            f"{worst:.2e} (layer 0: {worst0:.2e})")
    print(line)
    with open(os.environ.get("OCC_PARITY_LOG", os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "parity.log")), "a") as f:
        f.write(line + "\n")
    if os.environ.get("OCC_PARITY_REPORT_ONLY") != "1":
        assert worst0 < 3e-4, f"layer-0 attention-mask flips at non-negligible logits: {worst0:.2e}"
    for i in (0, DEC_LAYERS):
        assert_close(mask_cpu[i], ref["mask"][i], what=f"{tag} mask_pred[{i}]")
    assert_close(res["output_voxels"][0], ref["vox"], what=f"{tag} output_voxels {tuple(ref['vox'].shape)}")
    lab_ref = ref["vox"].argmax(1)
    agree = float((res["output_labels"].long().cpu() == lab_ref).double().mean())
    line = f"[parity] {tag} label agreement {100 * agree:.4f} %"
    print(line)
    with open(os.environ.get("OCC_PARITY_LOG", os.path.join(os.path.dirname(__file__), "..", "gpurun_out", "parity.log")), "a") as f:
        f.write(line + "\n")
    assert agree > 0.9999
