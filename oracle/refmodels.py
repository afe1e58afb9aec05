"""TEST INFRASTRUCTURE ONLY -- builds the *real* reference modules (imported verbatim from
/root/reference under oracle/shim.py) with the port's deterministic weights.  Works only where
/root/reference exists (the build container); used to validate oracle/port.py and to generate
tests/golden/.  Never imported on the GPU box."""
import torch

from . import port, shim

OCC = "projects.mmdet3d_plugin.occformer."
NORM_CFG = dict(type="GN", num_groups=32, requires_grad=True)


def ref_bev_pool_cpu(feats, coords, B, D, H, W):
    """CPU stand-in for the CUDA-only bev_pool_ext, driven by the reference's *own* pure-torch
    fallback semantics (QuickCumsum, ViewTransformerLSSBEVDepth.py:177-191): sort by rank,
    cumsum trick, scatter.  Independent of oracle/port.bev_pool on purpose."""
    vt = shim.load(OCC + "image2bev.ViewTransformerLSSBEVDepth")
    B_, D_, H_, W_ = int(B), int(D), int(H), int(W)
    ranks = coords[:, 0] * (W_ * D_ * B_) + coords[:, 1] * (D_ * B_) + coords[:, 2] * B_ + coords[:, 3]
    order = ranks.argsort(stable=True)
    x, gf, ranks = feats[order].double(), coords[order], ranks[order]
    x, gf = vt.cumsum_trick(x, gf, ranks)
    final = torch.zeros((B_, feats.shape[1], D_, H_, W_), dtype=torch.float64)
    final[gf[:, 3], :, gf[:, 2], gf[:, 0], gf[:, 1]] = x
    return final.float()


def build_block(cin, c, stride, layer_index, sd, prefix=""):
    shim.install()
    mod = shim.load(OCC + "backbones.dualpath_block")
    blk = mod.DualpathTransformerBlock(cin, c, stride=stride, norm_cfg=NORM_CFG, layer_index=layer_index)
    sub = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    blk.load_state_dict(sub, strict=True)
    return blk.eval()


def build_encoder(in_channels, block_inplanes, block_numbers, block_strides, out_indices, sd):
    shim.install()
    mod = shim.load(OCC + "backbones.occnet")
    enc = mod.OccupancyEncoder(in_channels=in_channels, num_stage=len(block_numbers),
                               block_numbers=list(block_numbers), block_inplanes=list(block_inplanes),
                               block_strides=list(block_strides), out_indices=tuple(out_indices),
                               norm_cfg=NORM_CFG, with_cp=True)
    enc.load_state_dict(sd, strict=True)
    return enc.eval()


def head_cfg(E, Q, K, num_layers, num_levels, ffn):
    return shim.to_cfg(dict(
        feat_channels=E, out_channels=E, num_queries=Q, num_occupancy_classes=K,
        num_transformer_feat_level=num_levels, pooling_attn_mask=True,
        positional_encoding=dict(type="SinePositionalEncoding3D", num_feats=E / 3, normalize=True),
        transformer_decoder=dict(
            type="DetrTransformerDecoder", return_intermediate=True, num_layers=num_layers,
            transformerlayers=dict(
                type="DetrTransformerDecoderLayer",
                attn_cfgs=dict(type="MultiheadAttention", embed_dims=E, num_heads=E // 32, attn_drop=0.0,
                               proj_drop=0.0, dropout_layer=None, batch_first=False),
                ffn_cfgs=dict(embed_dims=E, num_fcs=2, act_cfg=dict(type="ReLU", inplace=True),
                              ffn_drop=0.0, dropout_layer=None, add_identity=True),
                feedforward_channels=ffn,
                operation_order=("cross_attn", "norm", "self_attn", "norm", "ffn", "norm")),
            init_cfg=None),
        loss_cls=dict(type="CrossEntropyLoss", class_weight=[1.0] * K + [0.1]),
        loss_mask=dict(type="CrossEntropyLoss"), loss_dice=dict(type="DiceLoss"),
        point_cloud_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]))


def build_head(E, Q, K, num_layers, num_levels, ffn, sd, kitti=False):
    shim.install()
    shim.load(OCC + "mask2former.positional_encodings.positional_encoding")
    if kitti:
        cls = shim.load(OCC + "mask2former.mask2former_occ").Mask2FormerOccHead
    else:
        cls = shim.load(OCC + "mask2former.mask2former_nusc_occ").Mask2FormerNuscOccHead
    head = cls(**head_cfg(E, Q, K, num_layers, num_levels, ffn))
    missing, unexpected = head.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert not missing, missing
    return head.eval()


def build_view_transformer(grid_config, input_size, numC_input=64, numC_Trans=128, downsample=16):
    """ViewTransformerLiftSplatShootVoxel with the CPU bev_pool stand-in.  DepthNet needs mmcv DCN
    (un-vendored), so depth_net is never run here: callers test get_geometry / voxel_pooling and
    the lift (ViewTransformerLSSVoxel.py:110-119) fed with post-depth_net tensors."""
    shim.install()
    shim.BEV_POOL_IMPL["fn"] = ref_bev_pool_cpu
    mod = shim.load(OCC + "image2bev.ViewTransformerLSSVoxel")
    base = shim.load(OCC + "image2bev.ViewTransformerLSSBEVDepth")
    cls = mod.ViewTransformerLiftSplatShootVoxel
    vt = cls.__new__(cls)
    # run only the geometry part of the constructor chain (ViewTransformerLSSBEVDepth.py:64-99):
    base.ViewTransformerLiftSplatShoot.__init__(
        vt, grid_config=grid_config, data_config={"input_size": input_size}, numC_input=numC_input,
        numC_Trans=numC_Trans, downsample=downsample)
    vt.loss_depth_type = "bce"
    return vt.eval()


def ref_lift_and_pool(vt, depth_digit, img_feat, geom, B, N):
    """ViewTransformerLSSVoxel.forward :110-119 without depth_net (verbatim statements)."""
    depth_prob = vt.get_depth_dist(depth_digit)
    volume = depth_prob.unsqueeze(1) * img_feat.unsqueeze(2)
    H, W = depth_digit.shape[-2:]
    volume = volume.view(B, N, -1, vt.D, H, W)
    volume = volume.permute(0, 1, 3, 4, 5, 2)
    return vt.voxel_pooling(geom, volume), depth_prob


def neck_cfg(in_channels, strides, E, num_layers, num_heads, num_levels, num_points, ffn):
    return shim.to_cfg(dict(
        in_channels=list(in_channels), strides=list(strides), feat_channels=E, out_channels=E, num_outs=3,
        norm_cfg=dict(type="GN", num_groups=32), act_cfg=dict(type="ReLU"),
        encoder=dict(
            type="DetrTransformerEncoder", num_layers=num_layers,
            transformerlayers=dict(
                type="BaseTransformerLayer",
                attn_cfgs=dict(type="MultiScaleDeformableAttention3D", embed_dims=E, num_heads=num_heads,
                               num_levels=num_levels, num_points=num_points, im2col_step=64, dropout=0.0,
                               batch_first=False, norm_cfg=None, init_cfg=None),
                ffn_cfgs=dict(embed_dims=E), feedforward_channels=ffn, ffn_dropout=0.0,
                operation_order=("self_attn", "norm", "ffn", "norm")),
            init_cfg=None),
        positional_encoding=dict(type="SinePositionalEncoding3D", num_feats=E // 3, normalize=True)))


def build_neck(in_channels, strides, E, num_layers, num_heads, num_levels, num_points, ffn, sd):
    """The reference MSDeformAttnPixelDecoder3D (P/occformer/necks/multiscale_deformattn_3d.py) with the port's weights."""
    shim.install()
    shim.load(OCC + "mask2former.positional_encodings.positional_encoding")  # registers SinePositionalEncoding3D
    mod = shim.load(OCC + "necks.multiscale_deformattn_3d")
    neck = mod.MSDeformAttnPixelDecoder3D(**neck_cfg(in_channels, strides, E, num_layers, num_heads, num_levels,
                                                     num_points, ffn))
    neck.load_state_dict(sd, strict=True)
    return neck.eval()
