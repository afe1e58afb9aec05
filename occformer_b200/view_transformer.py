"""B200-native LSS lift-splat view transformer, drop-in for

    NECKS 'ViewTransformerLiftSplatShootVoxel'   projects/mmdet3d_plugin/occformer/image2bev/ViewTransformerLSSVoxel.py:12-121
    (geometry / frustum parts of its bases)      .../image2bev/ViewTransformerLSSBEVDepth.py:21-25,64-150
    mmdet3d.ops.bev_pool.bev_pool                mmdetection3d/mmdet3d/ops/bev_pool/bev_pool.py:83-97

``forward`` never materialises the (B,N,D,fH,fW,C) volume: depth softmax and NCHW->NHWC happen in a small
prologue and the pooling kernel multiplies depth and context on the fly.  ``voxel_pooling(geom, volume)`` and
``bev_pool(feats, coords, ...)`` keep the reference's materialised-input signatures for callers that use them.

DepthNet (mmcv DCN + ResNet BasicBlocks, ViewTransformerLSSBEVDepth.py:450-504) is NOT part of the replaced
hot path: ``depth_net`` is a plain attribute -- assign the reference's DepthNet when
running inside mmdetection3d; by default the base LSS 1x1 conv (ViewTransformerLSSBEVDepth.py:95) is built.
"""
import torch
import torch.nn as nn

from . import ops
from .registry import NECKS


def gen_dx_bx(xbound, ybound, zbound):
    """ViewTransformerLSSBEVDepth.py:21-25 (float32 torch.Tensor arithmetic, kept as is for bit-exactness)."""
    dx = torch.Tensor([row[2] for row in [xbound, ybound, zbound]])
    bx = torch.Tensor([row[0] + row[2] / 2.0 for row in [xbound, ybound, zbound]])
    nx = torch.Tensor([(row[1] - row[0]) / row[2] for row in [xbound, ybound, zbound]])
    return dx, bx, nx


def bev_pool(feats, coords, B, D, H, W):
    """Drop-in for mmdet3d.ops.bev_pool.bev_pool: feats (n,C) fp32 CUDA, coords (n,4) integer (x,y,z,b);
    B, D(=nz), H(=nx), W(=ny) ints or 0-d tensors.  Returns (B, C, D, H, W) (a permuted view of the
    channel-last grid; values identical to the reference's contiguous tensor).  Forward only."""
    assert feats.shape[0] == coords.shape[0]
    B, D, H, W = int(B), int(D), int(H), int(W)
    out = ops.bev_pool_channel_last(feats.float().contiguous(), coords.long().contiguous(), B, H, W, D)
    return out.permute(0, 4, 3, 1, 2)  # (B,X,Y,Z,C) -> (B,C,Z,X,Y)


@NECKS.register_module()
class ViewTransformerLiftSplatShootVoxel(nn.Module):
    def __init__(self, loss_depth_weight=1.0, grid_config=None, data_config=None, numC_input=512, numC_Trans=64,
                 downsample=16, point_cloud_range=None, loss_depth_type="bce", depth_net=None, **kwargs):
        super().__init__()
        if grid_config is None:
            grid_config = {"xbound": [-51.2, 51.2, 0.8], "ybound": [-51.2, 51.2, 0.8], "zbound": [-10.0, 10.0, 20.0],
                           "dbound": [1.0, 60.0, 1.0]}
        self.grid_config = grid_config
        dx, bx, nx = gen_dx_bx(grid_config["xbound"], grid_config["ybound"], grid_config["zbound"])
        self.dx = nn.Parameter(dx, requires_grad=False)
        self.bx = nn.Parameter(bx, requires_grad=False)
        self.nx = nn.Parameter(nx, requires_grad=False)
        self.data_config = data_config or {"input_size": (256, 704)}
        self.downsample = downsample
        self.frustum = self.create_frustum()
        self.D = self.frustum.shape[0]
        self.numC_input, self.numC_Trans = numC_input, numC_Trans
        self.depth_net = depth_net if depth_net is not None else nn.Conv2d(numC_input, self.D + numC_Trans, 1)
        self.loss_depth_weight = loss_depth_weight
        self.loss_depth_type = loss_depth_type
        self.cam_depth_range = grid_config["dbound"]
        self.point_cloud_range = point_cloud_range
        self._host = None

    def _host_params(self):
        # float copies of the tiny dx/bx/nx parameters for the by-value C ABI (one D2H, cached)
        if self._host is None:
            self._host = (self.dx.detach().cpu().tolist(), self.bx.detach().cpu().tolist(),
                          self.nx.detach().cpu().tolist())
        return self._host

    def _apply(self, fn, *a, **k):
        self._host = None
        return super()._apply(fn, *a, **k)

    def grid_size(self):
        _, _, nx = self._host_params()
        return int(nx[0]), int(nx[1]), int(nx[2])

    def create_frustum(self):
        """ViewTransformerLSSBEVDepth.py:104-115"""
        ogfH, ogfW = self.data_config["input_size"]
        fH, fW = ogfH // self.downsample, ogfW // self.downsample
        ds = torch.arange(*self.grid_config["dbound"], dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
        D = ds.shape[0]
        xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
        ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
        return nn.Parameter(torch.stack((xs, ys, ds), -1), requires_grad=False)

    def get_geometry(self, rots, trans, intrins, post_rots, post_trans, bda):
        """ViewTransformerLSSBEVDepth.py:117-150 as one fused kernel (occ_lss_geometry): (B,N,D,fH,fW,3) ego points."""
        if not trans.is_cuda:
            raise RuntimeError("occformer_b200: get_geometry runs on CUDA tensors only (no CPU fallback)")
        return ops.lss_geometry(self.frustum.data, rots, trans, intrins, post_rots, post_trans, bda)

    def get_depth_dist(self, x):
        return x.softmax(dim=1)

    @torch.no_grad()
    def voxel_pooling(self, geom_feats, x):
        """ViewTransformerLSSVoxel.py:77-100 with a materialised volume x (B,N,D,H,W,C).  -> (B,C,X,Y,Z) view."""
        B, N, D, H, W, C = x.shape
        dx, bx, nx = self._host_params()
        feats = x.reshape(B * N * D * H * W, C).float().contiguous()
        geom = geom_feats.reshape(-1, 3).float().contiguous()
        out = ops.voxel_pool_geom(feats, geom, B, dx, bx, nx, self.grid_size())
        return out.permute(0, 4, 1, 2, 3)

    @torch.no_grad()
    def lift_splat(self, depth_digit, img_feat, geom, B, N):
        """Fused lift-splat on post-DepthNet tensors: depth logits (B*N,D,fH,fW), context (B*N,C,fH,fW),
        geom (B,N,D,fH,fW,3).  Returns (channel-last grid (B,X,Y,Z,C), depth_prob)."""
        dx, bx, nx = self._host_params()
        prob, feat_cl = ops.lift_prologue(depth_digit.float().contiguous(), img_feat.float().contiguous())
        grid = ops.lift_splat(prob, feat_cl, geom.float().contiguous(), B, N, dx, bx, nx, self.grid_size())
        return grid, prob

    @torch.no_grad()
    def forward(self, input):
        """ViewTransformerLSSVoxel.py:102-121: returns (bev_feat (B,C,X,Y,Z), depth_prob (B*N,D,fH,fW))."""
        (x, rots, trans, intrins, post_rots, post_trans, bda, mlp_input) = input[:8]
        B, N, C, H, W = x.shape
        x = x.view(B * N, C, H, W)
        try:
            x = self.depth_net(x, mlp_input)
        except TypeError:
            x = self.depth_net(x)
        depth_digit = x[:, :self.D, ...]
        img_feat = x[:, self.D:self.D + self.numC_Trans, ...]
        geom = self.get_geometry(rots, trans, intrins, post_rots, post_trans, bda)
        grid, depth_prob = self.lift_splat(depth_digit, img_feat, geom, B, N)
        return grid.permute(0, 4, 1, 2, 3), depth_prob
