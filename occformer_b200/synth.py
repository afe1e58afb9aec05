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
