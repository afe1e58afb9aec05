"""DepthNet of the LSS view transformer -- SURVEY.md 8(f)3, a "next" row: the caller in front of the lift-splat hot path.

    DepthNet / ASPP / Mlp / SELayer      projects/mmdet3d_plugin/occformer/image2bev/ViewTransformerLSSBEVDepth.py:330-504
    BasicBlock                           mmdet 2.14.0 mmdet/models/backbones/resnet.py (conv1, bn1, conv2, bn2)
    DCN (mmcv DeformConv2dPack)          mmcv-full 1.4.0 mmcv/ops/deform_conv.py (weight, conv_offset.{weight,bias})

What is built here: the module tree with the reference's exact ``state_dict`` keys and shapes, so that a reference
checkpoint's ``img_view_transformer.depth_net.*`` entries load with ``strict=True``, and an inference forward on
library kernels (torch / cuDNN convolutions and ``grid_sample`` for the deformable convolution).  This is NOT one of the
hand-written sm_100a kernels of the hot path; it exists so the registered view transformer builds and runs
from the unchanged config.  BatchNorm runs in eval mode (running statistics), as in the reference's test path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class BasicBlock(nn.Module):
    """mmdet ResNet BasicBlock with stride 1, no downsample (DepthNet.depth_conv[0:3])."""

    def __init__(self, inplanes, planes):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=1, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return self.relu(out + x)


class _ASPPModule(nn.Module):
    def __init__(self, inplanes, planes, kernel_size, padding, dilation):
        super().__init__()
        self.atrous_conv = nn.Conv2d(inplanes, planes, kernel_size, stride=1, padding=padding, dilation=dilation, bias=False)
        self.bn = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU()

    def forward(self, x):
        return self.relu(self.bn(self.atrous_conv(x)))


class ASPP(nn.Module):
    """ViewTransformerLSSBEVDepth.py:349-414 (dropout 0.5 is the identity in eval)."""

    def __init__(self, inplanes, mid_channels=256):
        super().__init__()
        d = [1, 6, 12, 18]
        self.aspp1 = _ASPPModule(inplanes, mid_channels, 1, 0, d[0])
        self.aspp2 = _ASPPModule(inplanes, mid_channels, 3, d[1], d[1])
        self.aspp3 = _ASPPModule(inplanes, mid_channels, 3, d[2], d[2])
        self.aspp4 = _ASPPModule(inplanes, mid_channels, 3, d[3], d[3])
        self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Conv2d(inplanes, mid_channels, 1, bias=False),
                                             nn.BatchNorm2d(mid_channels), nn.ReLU())
        self.conv1 = nn.Conv2d(5 * mid_channels, mid_channels, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid_channels)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.5)

    def forward(self, x):
        x5 = F.interpolate(self.global_avg_pool(x), size=x.shape[2:], mode="bilinear", align_corners=True)
        y = torch.cat((self.aspp1(x), self.aspp2(x), self.aspp3(x), self.aspp4(x), x5), dim=1)
        return self.dropout(self.relu(self.bn1(self.conv1(y))))


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features, out_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.ReLU()
        self.drop1 = nn.Dropout(0.0)
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x):
        return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))


class SELayer(nn.Module):
    def __init__(self, channels):
        super().__init__()
        self.conv_reduce = nn.Conv2d(channels, channels, 1, bias=True)
        self.act1 = nn.ReLU()
        self.conv_expand = nn.Conv2d(channels, channels, 1, bias=True)
        self.gate = nn.Sigmoid()

    def forward(self, x, x_se):
        return x * self.gate(self.conv_expand(self.act1(self.conv_reduce(x_se))))


class DCN(nn.Module):
    """mmcv ``DeformConv2dPack`` (DCNv1): 3x3, stride 1, padding 1, ``groups`` weight groups, one deformable group.
    Parameters: ``weight`` (out, in/groups, 3, 3) without bias and ``conv_offset`` = Conv2d(in, 18, 3, padding 1) whose
    output channel 2k / 2k+1 is the (dy, dx) offset of kernel point k.  Forward: bilinear sampling with zeros outside the
    map (``grid_sample``), then the grouped contraction."""

    def __init__(self, in_channels, out_channels, kernel_size=3, padding=1, groups=4, **_):
        super().__init__()
        assert kernel_size == 3 and padding == 1
        self.groups = groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, 3, 3))
        nn.init.kaiming_uniform_(self.weight, nonlinearity="relu")
        self.conv_offset = nn.Conv2d(in_channels, 18, 3, stride=1, padding=1, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x):
        B, C, H, W = x.shape
        off = self.conv_offset(x)
        ys = torch.arange(H, device=x.device, dtype=x.dtype).view(1, H, 1)
        xs = torch.arange(W, device=x.device, dtype=x.dtype).view(1, 1, W)
        cols = []
        for k in range(9):
            py = ys + (k // 3 - 1) + off[:, 2 * k]
            px = xs + (k % 3 - 1) + off[:, 2 * k + 1]
            grid = torch.stack((2.0 * px / max(W - 1, 1) - 1.0, 2.0 * py / max(H - 1, 1) - 1.0), dim=-1)
            cols.append(F.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=True))
        col = torch.stack(cols, dim=2)  # (B, C, 9, H, W)
        g = self.groups
        col = col.view(B, g, C // g, 9, H, W)
        w = self.weight.view(g, -1, C // g, 9)  # (g, out/g, in/g, 9)
        return torch.einsum("bgckhw,gock->bgohw", col, w).reshape(B, -1, H, W)


class DepthNet(nn.Module):
    """ViewTransformerLSSBEVDepth.py:450-504: camera-aware depth / context heads.  forward(x (B*N, Cin, fH, fW),
    mlp_input (B, N, cam_channels)) -> (B*N, D + C_ctx, fH, fW) = depth logits | context features."""

    def __init__(self, in_channels, mid_channels, context_channels, depth_channels, cam_channels=27):
        super().__init__()
        self.reduce_conv = nn.Sequential(nn.Conv2d(in_channels, mid_channels, 3, stride=1, padding=1),
                                         nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True))
        self.context_conv = nn.Conv2d(mid_channels, context_channels, 1)
        self.bn = nn.BatchNorm1d(cam_channels)
        self.depth_mlp = Mlp(cam_channels, mid_channels, mid_channels)
        self.depth_se = SELayer(mid_channels)
        self.context_mlp = Mlp(cam_channels, mid_channels, mid_channels)
        self.context_se = SELayer(mid_channels)
        self.depth_conv = nn.Sequential(BasicBlock(mid_channels, mid_channels), BasicBlock(mid_channels, mid_channels),
                                        BasicBlock(mid_channels, mid_channels), ASPP(mid_channels, mid_channels),
                                        DCN(mid_channels, mid_channels, 3, padding=1, groups=4),
                                        nn.Conv2d(mid_channels, depth_channels, 1))

    def forward(self, x, mlp_input):
        mlp_input = self.bn(mlp_input.reshape(-1, mlp_input.shape[-1]))
        x = self.reduce_conv(x)
        context = self.context_conv(self.context_se(x, self.context_mlp(mlp_input)[..., None, None]))
        depth = self.depth_conv(self.depth_se(x, self.depth_mlp(mlp_input)[..., None, None]))
        return torch.cat([depth, context], dim=1)
