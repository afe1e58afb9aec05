"""Deterministic synthetic inputs for the hot path: nuScenes-like fixed camera
constants, seeded CPU generation.  Shared by tests/ and bench.py so that the oracle and the CUDA
path always see identical tensors.  No dataset, no checkpoint: everything is synthetic."""
import math

import torch

GRIDS = {
    # name: (xbound, ybound, zbound)            -> X x Y x Z
    "pr1": ([-20.0, 20.0, 0.8], [-20.0, 20.0, 0.8], [-2.0, 4.4, 0.8]),          # 50 x 50 x 8
    "nusc_ref": ([-51.2, 51.2, 0.8], [-51.2, 51.2, 0.8], [-5.0, 3.0, 0.5]),     # 128 x 128 x 16
    "nusc_200": ([-40.0, 40.0, 0.4], [-40.0, 40.0, 0.4], [-1.0, 5.4, 0.4]),     # 200 x 200 x 16
    "kitti": ([0.0, 51.2, 0.4], [-25.6, 25.6, 0.4], [-2.0, 4.4, 0.4]),          # 128 x 128 x 16 (occformer_kitti.py:22-46)
}
DBOUND = [2.0, 58.0, 0.5]


def grid_config(name):
    xb, yb, zb = GRIDS[name]
    return {"xbound": xb, "ybound": yb, "zbound": zb, "dbound": DBOUND}


def _yaw(deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


CAM2EGO_AXES = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])


# the workloads of BASELINE.json `configs` (shapes: SURVEY.md section 8 legend): cameras, input, lift grid, output grid, classes
WORKLOADS = {
    "nusc_200": dict(name="nusc_r50_6cam_256x704_to_200x200x16", cams=6, input_size=(256, 704), grid="nusc_200",
                     occ=[200, 200, 16], pc=[-40.0, -40.0, -1.0, 40.0, 40.0, 5.4], classes=17, head="Mask2FormerNuscOccHead"),
    "nusc_ref": dict(name="nusc_r50_6cam_256x704_to_128x128x16_occ256x256x32", cams=6, input_size=(256, 704), grid="nusc_ref",
                     occ=[256, 256, 32], pc=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], classes=17, head="Mask2FormerNuscOccHead"),
    "kitti": dict(name="semantickitti_1cam_384x1280_to_128x128x16_occ256x256x32", cams=1, input_size=(384, 1280), grid="kitti",
                  occ=[256, 256, 32], pc=[0.0, -25.6, -2.0, 51.2, 25.6, 4.4], classes=20, head="Mask2FormerOccHead"),
    "nusc_r101": dict(name="nusc_r101_6cam_896x1600_to_200x200x16", cams=6, input_size=(896, 1600), grid="nusc_200",
                      occ=[200, 200, 16], pc=[-40.0, -40.0, -1.0, 40.0, 40.0, 5.4], classes=17, head="Mask2FormerNuscOccHead"),
}


def workload_cameras(wl, B=1):
    w = WORKLOADS[wl]
    return kitti_camera(B, w["input_size"]) if wl == "kitti" else nusc_cameras(B, w["cams"], w["input_size"])


def nusc_cameras(B=1, N=6, input_size=(256, 704)):
    """Six cameras, yaw {+55,0,-55,+110,180,-110} deg; intrinsics of a 1600x900 nuScenes camera;
    post_rot/post_tran of the reference's test-time resize+crop (loading_nusc_imgs.py:85-93)."""
    yaws = [55.0, 0.0, -55.0, 110.0, 180.0, -110.0][:N]
    H, W = input_size
    resize = W / 1600.0
    crop_h = int(900 * resize) - H
    rots = torch.stack([_yaw(y) @ CAM2EGO_AXES for y in yaws])
    trans = torch.tensor([[1.5 * math.cos(math.radians(y)), 1.5 * math.sin(math.radians(y)), 1.5] for y in yaws])
    K = torch.tensor([[1266.4, 0.0, 816.3], [0.0, 1266.4, 491.5], [0.0, 0.0, 1.0]])
    intrins = K.expand(N, 3, 3)
    post_rots = torch.diag(torch.tensor([resize, resize, 1.0])).expand(N, 3, 3)
    post_trans = torch.tensor([0.0, -float(crop_h), 0.0]).expand(N, 3)
    bda = torch.eye(3)

    def rep(t):
        return t.unsqueeze(0).repeat(B, *([1] * t.dim())).contiguous()

    return dict(rots=rep(rots), trans=rep(trans), intrins=rep(intrins), post_rots=rep(post_rots),
                post_trans=rep(post_trans), bda=rep(bda))


def kitti_camera(B=1, input_size=(384, 1280)):
    """SemanticKITTI-like single left camera (BASELINE.json configs[1]): 4x4 P2 intrinsics with the shift column
    (semantic_kitti_lss_dataset.py:62-64 hands a 4x4 matrix over), camera -> LiDAR axes, 4x4 homogeneous bda."""
    H, W = input_size
    resize = W / 1226.0
    rots = (_yaw(0.0) @ CAM2EGO_AXES).view(1, 1, 3, 3).repeat(B, 1, 1, 1)
    trans = torch.tensor([0.27, 0.0, -0.08]).view(1, 1, 3).repeat(B, 1, 1)
    K = torch.eye(4)
    K[:3, :4] = torch.tensor([[707.09, 0.0, 604.08, 45.76], [0.0, 707.09, 180.51, -0.35], [0.0, 0.0, 1.0, 0.005]])
    post_rots = torch.diag(torch.tensor([resize, resize, 1.0])).view(1, 1, 3, 3).repeat(B, 1, 1, 1)
    post_trans = torch.tensor([0.0, -float(int(370 * resize) - H), 0.0]).view(1, 1, 3).repeat(B, 1, 1)
    return dict(rots=rots.contiguous(), trans=trans.contiguous(), intrins=K.view(1, 1, 4, 4).repeat(B, 1, 1, 1).contiguous(),
                post_rots=post_rots.contiguous(), post_trans=post_trans.contiguous(),
                bda=torch.eye(4).view(1, 4, 4).repeat(B, 1, 1).contiguous())


def pr1_camera(B=1):
    """One front camera, 128x128, fx=fy=100, cx=cy=64 (BASELINE.json configs[0])."""
    rots = (_yaw(0.0) @ CAM2EGO_AXES).view(1, 1, 3, 3).repeat(B, 1, 1, 1)
    trans = torch.tensor([1.5, 0.0, 1.5]).view(1, 1, 3).repeat(B, 1, 1)
    K = torch.tensor([[100.0, 0.0, 64.0], [0.0, 100.0, 64.0], [0.0, 0.0, 1.0]])
    return dict(rots=rots, trans=trans, intrins=K.view(1, 1, 3, 3).repeat(B, 1, 1, 1),
                post_rots=torch.eye(3).view(1, 1, 3, 3).repeat(B, 1, 1, 1),
                post_trans=torch.zeros(B, 1, 3), bda=torch.eye(3).view(1, 3, 3).repeat(B, 1, 1))


def lift_inputs(B, N, D, fH, fW, C, seed=0):
    """Post-DepthNet tensors: depth logits ~ 3*N(0,1) (B*N,D,fH,fW), context ~ N(0,1) (B*N,C,fH,fW)."""
    g = torch.Generator().manual_seed(seed)
    depth_digit = 3.0 * torch.randn(B * N, D, fH, fW, generator=g)
    img_feat = torch.randn(B * N, C, fH, fW, generator=g)
    return depth_digit, img_feat


def encoder_input(B, C, X, Y, Z, seed=0):
    g = torch.Generator().manual_seed(seed)
    return 0.5 * torch.randn(B, C, X, Y, Z, generator=g)


def head_inputs(B, E, sizes, seed=0):
    """voxel_feats list, high -> low resolution: [mask_features, level2, level1, level0]."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, E, *s, generator=g) * (0.5 if i == 0 else 1.0) for i, s in enumerate(sizes)]


def lidar_points(n, pc_range, seed=0):
    g = torch.Generator().manual_seed(seed)
    lo = torch.tensor(pc_range[:3])
    hi = torch.tensor(pc_range[3:])
    pts = lo + (hi - lo) * (torch.rand(n, 3, generator=g) * 1.1 - 0.05)  # a few out of range
    return torch.cat([pts, torch.zeros(n, 2)], dim=1)


# =============================================================================
# deterministic synthetic weights (same keys / shapes as the reference modules)
# =============================================================================
def rel_position_index(ws=7):
    """WindowMSA.relative_position_index (window_attention.py:57-61, 109-113): double_step_seq + flip."""
    seq1 = torch.arange(0, (2 * ws - 1) * ws, 2 * ws - 1)
    seq2 = torch.arange(0, ws, 1)
    c = (seq1[:, None] + seq2[None, :]).reshape(1, -1)
    return (c + c.T).flip(1).contiguous()



def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g) * std


def make_block_state(cin, c, stride, g, prefix=""):
    """state_dict of one DualpathTransformerBlock, seeded generator ``g``.
    Zero-initialised reference params are perturbed so that they matter."""
    sd = {}
    p = prefix

    def norm(name, n):
        sd[p + name + "weight"] = 1 + 0.1 * _randn(g, n)
        sd[p + name + "bias"] = 0.1 * _randn(g, n)

    def lin(name, o, i):
        sd[p + name + "weight"] = _randn(g, o, i, std=i ** -0.5)
        sd[p + name + "bias"] = 0.1 * _randn(g, o)

    if stride > 1:
        sd[p + "downsample.0.weight"] = _randn(g, c, cin, 1, 1, 1, std=cin ** -0.5)
        norm("downsample.1.", c)
    sd[p + "input_conv.0.weight"] = _randn(g, c, cin, 3, 3, 3, std=(27 * cin) ** -0.5)
    norm("input_conv.1.", c)
    be = "bev_encoder."
    norm(be + "norm1.", c)
    sd[p + be + "attn.w_msa.relative_position_bias_table"] = _randn(g, 169, c // 32, std=0.5)
    sd[p + be + "attn.w_msa.relative_position_index"] = rel_position_index(7)
    lin(be + "attn.w_msa.qkv.", 3 * c, c)
    lin(be + "attn.w_msa.proj.", c, c)
    norm(be + "norm2.", c)
    lin(be + "ffn.layers.0.0.", c, c)
    lin(be + "ffn.layers.1.", c, c)
    ch = c // 4
    sd[p + "aspp.input_conv.0.weight"] = _randn(g, ch, c, 1, 1, std=c ** -0.5)
    norm("aspp.input_conv.1.", ch)
    sd[p + "aspp.aspp.aspp1.atrous_conv.weight"] = _randn(g, ch, ch, 1, 1, std=ch ** -0.5)
    norm("aspp.aspp.aspp1.bn.", ch)
    for i in (2, 3, 4):
        sd[p + f"aspp.aspp.aspp{i}.atrous_conv.weight"] = _randn(g, ch, ch, 3, 3, std=(9 * ch) ** -0.5)
        norm(f"aspp.aspp.aspp{i}.bn.", ch)
    sd[p + "aspp.aspp.global_avg_pool.1.weight"] = _randn(g, ch, ch, 1, 1, std=ch ** -0.5)
    norm("aspp.aspp.global_avg_pool.2.", ch)
    sd[p + "aspp.aspp.conv1.weight"] = _randn(g, ch, 5 * ch, 1, 1, std=(5 * ch) ** -0.5)
    norm("aspp.aspp.bn1.", ch)
    sd[p + "aspp.output_conv.0.weight"] = _randn(g, c, ch, 1, 1, std=ch ** -0.5)
    norm("aspp.output_conv.1.", c)
    sd[p + "combine_coeff.weight"] = _randn(g, 1, c, 1, 1, 1, std=c ** -0.5)
    sd[p + "combine_coeff.bias"] = 0.1 * _randn(g, 1)
    return sd


def make_encoder_state(in_channels, block_inplanes, block_numbers, block_strides, seed=0, prefix=""):
    g = torch.Generator().manual_seed(seed)
    sd = {}
    cin = in_channels
    for s, (c, nb, st) in enumerate(zip(block_inplanes, block_numbers, block_strides)):
        for b in range(nb):
            sd.update(make_block_state(cin, c, st if b == 0 else 1, g, f"{prefix}layers.{s}.{b}."))
            cin = c
    return sd


def make_head_state(E=192, Q=100, K=17, num_layers=9, num_levels=3, ffn=None, seed=0, prefix=""):
    """state_dict of Mask2FormerNuscOccHead (inference params)."""
    g = torch.Generator().manual_seed(seed)
    ffn = ffn or 8 * E
    sd = {}
    p = prefix

    def norm(name):
        sd[p + name + "weight"] = 1 + 0.1 * _randn(g, E)
        sd[p + name + "bias"] = 0.1 * _randn(g, E)

    def lin(name, o, i, wname="weight", bname="bias"):
        sd[p + name + wname] = _randn(g, o, i, std=i ** -0.5)
        sd[p + name + bname] = 0.1 * _randn(g, o)

    sd[p + "query_embed.weight"] = _randn(g, Q, E)
    sd[p + "query_feat.weight"] = _randn(g, Q, E)
    sd[p + "level_embed.weight"] = _randn(g, num_levels, E)
    lin("cls_embed.", K + 1, E)
    for i in (0, 2, 4):
        lin(f"mask_embed.{i}.", E, E)
    for l in range(num_layers):
        lp = f"transformer_decoder.layers.{l}."
        for a in (0, 1):
            lin(lp + f"attentions.{a}.attn.", 3 * E, E, "in_proj_weight", "in_proj_bias")
            lin(lp + f"attentions.{a}.attn.out_proj.", E, E)
        lin(lp + "ffns.0.layers.0.0.", ffn, E)
        lin(lp + "ffns.0.layers.1.", E, ffn)
        for n in (0, 1, 2):
            norm(lp + f"norms.{n}.")
    norm("transformer_decoder.post_norm.")
    return sd


def make_neck_state(in_channels, E, num_layers, num_heads, num_levels=3, num_points=4, ffn=None, seed=0):
    """Deterministic weights with the reference's state_dict keys.  Unlike the reference's init (zero sampling-offset
    and attention-weight matrices), every matrix is non-trivial so that a parity test moves the sampling points."""
    g = torch.Generator().manual_seed(seed)
    ffn = ffn or 4 * E
    sd = {}

    def rn(*shape, std=1.0):
        return torch.randn(*shape, generator=g) * std

    def gn(name, n):
        sd[name + "weight"] = 1.0 + 0.1 * rn(n)
        sd[name + "bias"] = 0.1 * rn(n)

    nin = len(in_channels)
    for i in range(num_levels):
        cin = in_channels[nin - i - 1]
        sd[f"input_convs.{i}.conv.weight"] = rn(E, cin, 1, 1, 1, std=cin ** -0.5)
        sd[f"input_convs.{i}.conv.bias"] = 0.1 * rn(E)
        gn(f"input_convs.{i}.gn.", E)
    for l in range(num_layers):
        p = f"encoder.layers.{l}."
        n_off = num_heads * num_levels * num_points
        sd[p + "attentions.0.sampling_offsets.weight"] = rn(3 * n_off, E, std=0.5 * E ** -0.5)
        sd[p + "attentions.0.sampling_offsets.bias"] = rn(3 * n_off, std=1.0)
        sd[p + "attentions.0.attention_weights.weight"] = rn(n_off, E, std=E ** -0.5)
        sd[p + "attentions.0.attention_weights.bias"] = 0.1 * rn(n_off)
        for nm in ("value_proj", "output_proj"):
            sd[p + f"attentions.0.{nm}.weight"] = rn(E, E, std=E ** -0.5)
            sd[p + f"attentions.0.{nm}.bias"] = 0.1 * rn(E)
        sd[p + "ffns.0.layers.0.0.weight"] = rn(ffn, E, std=E ** -0.5)
        sd[p + "ffns.0.layers.0.0.bias"] = 0.1 * rn(ffn)
        sd[p + "ffns.0.layers.1.weight"] = rn(E, ffn, std=ffn ** -0.5)
        sd[p + "ffns.0.layers.1.bias"] = 0.1 * rn(E)
        gn(p + "norms.0.", E)
        gn(p + "norms.1.", E)
    sd["level_encoding.weight"] = rn(num_levels, E)
    for i in range(nin - num_levels):
        sd[f"lateral_convs.{i}.conv.weight"] = rn(E, in_channels[i], 1, 1, 1, std=in_channels[i] ** -0.5)
        gn(f"lateral_convs.{i}.gn.", E)
        sd[f"output_convs.{i}.conv.weight"] = rn(E, E, 3, 3, 3, std=(27 * E) ** -0.5)
        gn(f"output_convs.{i}.gn.", E)
    sd["mask_feature.weight"] = rn(E, E, 1, 1, 1, std=E ** -0.5)
    sd["mask_feature.bias"] = 0.1 * rn(E)
    return sd
