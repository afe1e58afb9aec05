"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

A minimal stand-in for the un-vendored third-party packages the reference's
hot-path files import (mmcv-full==1.4.0, mmdet==2.14.0, mmdet3d 0.17.1 --
pinned in /root/reference/docs/install.md:15-24 and gated in
mmdetection3d/mmdet3d/__init__.py:22-49).  With this shim on ``sys.modules``
the reference files under /root/reference/projects/mmdet3d_plugin import
*verbatim* and run on CPU.  Only the symbols those files touch are provided;
their arithmetic is a thin layer over torch.nn, restated here from the published
mmcv 1.4.0 sources (mmcv/cnn/bricks/transformer.py, norm.py, conv_module.py)
and mmdet 2.14.0 (mmdet/models/utils/transformer.py).

The shim exists so that (1) oracle/port.py (the CPU restatement that travels to
the GPU box) can be validated against the real reference code in this container
and (2) tests/golden/ fixtures can be generated from the real reference.
/root/reference does not exist on the GPU box; nothing at GPU-test/bench time
imports this file.
"""
import copy
import importlib
import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("OCC_REFERENCE_ROOT", "/root/reference")
PLUGIN = os.path.join(REFERENCE_ROOT, "projects", "mmdet3d_plugin")


# --------------------------------------------------------------------------- utils
class ConfigDict(dict):
    """Attribute-accessible dict (mmcv.utils.ConfigDict semantics we need)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def to_cfg(d):
    if isinstance(d, dict):
        return ConfigDict({k: to_cfg(v) for k, v in d.items()})
    if isinstance(d, (list, tuple)):
        return type(d)(to_cfg(v) for v in d)
    return d


class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self.name = name
        self._module_dict = {}
        self.build_func = build_func or build_from_cfg

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self._module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def build(self, cfg, *a, **k):
        return self.build_func(cfg, self, *a, **k) if self.build_func is not build_from_cfg \
            else build_from_cfg(cfg, self, *a, **k)


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop("type")
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError(f"{t} is not in the {registry.name} registry")
    return cls(**args)


def _noop_decorator_factory(*a, **k):
    def deco(f):
        return f
    return deco


def _noop(*a, **k):
    return None


# --------------------------------------------------------------------------- mmcv.runner
class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self._is_init = False
        self.init_cfg = copy.deepcopy(init_cfg)

    def init_weights(self):
        pass


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


# --------------------------------------------------------------------------- mmcv.cnn
NORM_LAYERS = {
    "BN": (nn.BatchNorm2d, "bn"), "BN1d": (nn.BatchNorm1d, "bn"), "BN2d": (nn.BatchNorm2d, "bn"),
    "BN3d": (nn.BatchNorm3d, "bn"), "GN": (nn.GroupNorm, "gn"), "LN": (nn.LayerNorm, "ln"),
    "SyncBN": (nn.BatchNorm2d, "bn"),
}


def build_norm_layer(cfg, num_features, postfix=""):
    cfg_ = dict(cfg)
    t = cfg_.pop("type")
    layer_cls, abbr = NORM_LAYERS[t]
    requires_grad = cfg_.pop("requires_grad", True)
    cfg_.setdefault("eps", 1e-5)
    if t == "GN":
        layer = layer_cls(num_channels=num_features, **cfg_)
    else:
        layer = layer_cls(num_features, **cfg_)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return abbr + str(postfix), layer


class DeformConv2dPack(nn.Module):
    """mmcv-full 1.4.0 ``mmcv/ops/deform_conv.py`` DeformConv2dPack ('DCN' in CONV_LAYERS), restated on CPU (the op is a
    CUDA-only extension in mmcv): ``weight`` (out, in/groups, kh, kw) without bias, ``conv_offset`` = Conv2d(in,
    deform_groups*2*kh*kw, k, stride, padding, dilation, bias=True) zero-initialised; forward = deform_conv2d(x, offset,
    weight) where offset channel 2k / 2k+1 of a deformable group is the (dy, dx) of kernel point k, sampling is bilinear
    with zero contribution from out-of-map corners (deform_conv_cuda_kernel.cuh dmcn_im2col_bilinear semantics of v1:
    points with h <= -1 || h >= H || w <= -1 || w >= W give 0).  Explicit corner gathers -- deliberately NOT grid_sample,
    so that it is an independent check of occformer_b200/depthnet.py."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, deform_groups=1,
                 bias=False, im2col_step=32, **kw):
        super().__init__()
        assert not bias and deform_groups == 1 and stride == 1
        self.k, self.padding, self.dilation, self.groups = kernel_size, padding, dilation, groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, kernel_size, kernel_size))
        nn.init.kaiming_uniform_(self.weight, nonlinearity="relu")
        self.conv_offset = nn.Conv2d(in_channels, 2 * kernel_size * kernel_size, kernel_size, stride=stride,
                                     padding=padding, dilation=dilation, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward(self, x):
        B, C, H, W = x.shape
        k, pad, dil = self.k, self.padding, self.dilation
        off = self.conv_offset(x)
        ys = torch.arange(H, dtype=x.dtype).view(1, H, 1)
        xs = torch.arange(W, dtype=x.dtype).view(1, 1, W)
        xf = x.reshape(B, C, H * W)
        cols = []
        for kk in range(k * k):
            py = ys - pad + (kk // k) * dil + off[:, 2 * kk]
            px = xs - pad + (kk % k) * dil + off[:, 2 * kk + 1]
            y0, x0 = torch.floor(py), torch.floor(px)
            val = torch.zeros(B, C, H, W, dtype=x.dtype)
            for dy in (0, 1):
                for dx in (0, 1):
                    yy, xx = y0 + dy, x0 + dx
                    w = (1 - (py - yy).abs()) * (1 - (px - xx).abs())
                    ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
                    idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).long().view(B, 1, H * W).expand(B, C, H * W)
                    val = val + torch.gather(xf, 2, idx).view(B, C, H, W) * (w * ok).unsqueeze(1)
            cols.append(val)
        col = torch.stack(cols, dim=2)  # (B, C, k*k, H, W)
        g = self.groups
        col = col.view(B, g, (C // g) * k * k, H * W)
        w = self.weight.view(g, -1, (C // g) * k * k)
        return torch.einsum("bgkp,gok->bgop", col, w).reshape(B, -1, H, W)


class BasicBlock(nn.Module):
    """mmdet 2.14.0 ``mmdet/models/backbones/resnet.py`` BasicBlock as DepthNet uses it (stride 1, no downsample, BN):
    conv1 -> bn1 -> relu -> conv2 -> bn2 -> (+identity) -> relu; norm layers are registered under 'bn1' / 'bn2'."""

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None, style="pytorch", with_cp=False, conv_cfg=None,
                 norm_cfg=dict(type="BN"), dcn=None, plugins=None, init_cfg=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.add_module("bn1", nn.BatchNorm2d(planes))
        self.conv2 = nn.Conv2d(planes, planes, 3, padding=1, bias=False)
        self.add_module("bn2", nn.BatchNorm2d(planes))
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample

    def forward(self, x):
        identity = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            identity = self.downsample(x)
        return self.relu(out + identity)


CONV_LAYERS = {"Conv1d": nn.Conv1d, "Conv2d": nn.Conv2d, "Conv3d": nn.Conv3d, "Conv": nn.Conv2d, "DCN": DeformConv2dPack}


def build_conv_layer(cfg, *args, **kwargs):
    if cfg is None:
        cfg_ = dict(type="Conv2d")
    else:
        cfg_ = dict(cfg)
    t = cfg_.pop("type")
    return CONV_LAYERS[t](*args, **kwargs, **cfg_)


ACT_LAYERS = {"ReLU": nn.ReLU, "GELU": nn.GELU, "Sigmoid": nn.Sigmoid, "LeakyReLU": nn.LeakyReLU}


def build_activation_layer(cfg):
    cfg_ = dict(cfg)
    t = cfg_.pop("type")
    if t == "GELU":
        cfg_.pop("inplace", None)
    return ACT_LAYERS[t](**cfg_)


class ConvModule(nn.Module):
    """conv -> norm -> act; bias='auto' => bias = not with_norm (mmcv 1.4.0 conv_module.py)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"),
                 inplace=True, with_spectral_norm=False, padding_mode="zeros",
                 order=("conv", "norm", "act")):
        super().__init__()
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride,
                                     padding=padding, dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            act_cfg_ = dict(act_cfg)
            if act_cfg_["type"] not in ("GELU",):
                act_cfg_.setdefault("inplace", inplace)
            self.activate = build_activation_layer(act_cfg_)

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.with_norm else None

    def forward(self, x, activate=True, norm=True):
        x = self.conv(x)
        if norm and self.with_norm:
            x = self.norm(x)
        if activate and self.with_activation:
            x = self.activate(x)
        return x


# --------------------------------------------------------------------------- mmcv.cnn.bricks.transformer
ATTENTION = Registry("attention")
POSITIONAL_ENCODING = Registry("position encoding")
TRANSFORMER_LAYER = Registry("transformerLayer")
TRANSFORMER_LAYER_SEQUENCE = Registry("transformer-layers sequence")
FEEDFORWARD_NETWORK = Registry("feed-forward Network")


def build_dropout(cfg, default_args=None):
    """DropPath / Dropout: identity in eval (parity is defined in .eval())."""
    if cfg is None:
        return nn.Identity()
    t = cfg.get("type")
    p = cfg.get("drop_prob", cfg.get("p", 0.0))
    if t == "DropPath":
        return DropPath(p)
    return nn.Dropout(p)


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.1):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        r = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        return x.div(keep) * r.floor()


def build_positional_encoding(cfg, default_args=None):
    return build_from_cfg(cfg, POSITIONAL_ENCODING, default_args)


def build_attention(cfg, default_args=None):
    return build_from_cfg(cfg, ATTENTION, default_args)


def build_feedforward_network(cfg, default_args=None):
    return build_from_cfg(cfg, FEEDFORWARD_NETWORK, default_args)


def build_transformer_layer(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER, default_args)


def build_transformer_layer_sequence(cfg, default_args=None):
    return build_from_cfg(cfg, TRANSFORMER_LAYER_SEQUENCE, default_args)


@ATTENTION.register_module()
class MultiheadAttention(BaseModule):
    def __init__(self, embed_dims, num_heads, attn_drop=0.0, proj_drop=0.0,
                 dropout_layer=dict(type="Dropout", drop_prob=0.0), init_cfg=None,
                 batch_first=False, **kwargs):
        super().__init__(init_cfg)
        if "dropout" in kwargs:
            attn_drop = kwargs["dropout"]
            dropout_layer = dict(dropout_layer or dict(type="Dropout"))
            dropout_layer["drop_prob"] = kwargs.pop("dropout")
        self.embed_dims = embed_dims
        self.num_heads = num_heads
        self.batch_first = batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None,
                attn_mask=None, key_padding_mask=None, **kwargs):
        if key is None:
            key = query
        if value is None:
            value = key
        if identity is None:
            identity = query
        if key_pos is None:
            if query_pos is not None:
                if query_pos.shape == key.shape:
                    key_pos = query_pos
                else:
                    warnings.warn("position encoding of key is missing in MultiheadAttention.")
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        if self.batch_first:
            query, key, value = query.transpose(0, 1), key.transpose(0, 1), value.transpose(0, 1)
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask,
                        key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


@FEEDFORWARD_NETWORK.register_module()
class FFN(BaseModule):
    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2,
                 act_cfg=dict(type="ReLU", inplace=True), ffn_drop=0.0, dropout_layer=None,
                 add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        assert num_fcs >= 2
        self.embed_dims = embed_dims
        self.feedforward_channels = feedforward_channels
        self.num_fcs = num_fcs
        self.activate = build_activation_layer(act_cfg)
        layers = []
        in_channels = embed_dims
        for _ in range(num_fcs - 1):
            layers.append(Sequential(nn.Linear(in_channels, feedforward_channels), self.activate,
                                     nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = Sequential(*layers)
        self.dropout_layer = build_dropout(dropout_layer) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        if identity is None:
            identity = x
        return identity + self.dropout_layer(out)


@TRANSFORMER_LAYER.register_module()
class BaseTransformerLayer(BaseModule):
    def __init__(self, attn_cfgs=None, ffn_cfgs=dict(type="FFN", embed_dims=256,
                 feedforward_channels=1024, num_fcs=2, ffn_drop=0.0,
                 act_cfg=dict(type="ReLU", inplace=True)), operation_order=None,
                 norm_cfg=dict(type="LN"), init_cfg=None, batch_first=False, **kwargs):
        deprecated = dict(feedforward_channels="feedforward_channels", ffn_dropout="ffn_drop",
                          ffn_num_fcs="num_fcs")
        ffn_cfgs = copy.deepcopy(dict(ffn_cfgs))
        for ori, new in deprecated.items():
            if ori in kwargs:
                ffn_cfgs[new] = kwargs[ori]
        super().__init__(init_cfg)
        self.batch_first = batch_first
        num_attn = operation_order.count("self_attn") + operation_order.count("cross_attn")
        if isinstance(attn_cfgs, dict):
            attn_cfgs = [copy.deepcopy(attn_cfgs) for _ in range(num_attn)]
        self.num_attn = num_attn
        self.operation_order = operation_order
        self.norm_cfg = norm_cfg
        self.pre_norm = operation_order[0] == "norm"
        self.attentions = ModuleList()
        index = 0
        for name in operation_order:
            if name in ("self_attn", "cross_attn"):
                cfg = dict(attn_cfgs[index])
                if "batch_first" in cfg:
                    assert self.batch_first == cfg["batch_first"]
                else:
                    cfg["batch_first"] = self.batch_first
                attention = build_attention(cfg)
                attention.operation_name = name
                self.attentions.append(attention)
                index += 1
        self.embed_dims = self.attentions[0].embed_dims
        self.ffns = ModuleList()
        num_ffns = operation_order.count("ffn")
        if isinstance(ffn_cfgs, dict):
            ffn_cfgs = [copy.deepcopy(ffn_cfgs) for _ in range(num_ffns)]
        for i in range(num_ffns):
            c = dict(ffn_cfgs[i])
            c.setdefault("type", "FFN")
            if "embed_dims" not in c:
                c["embed_dims"] = self.embed_dims
            self.ffns.append(build_feedforward_network(c))
        self.norms = ModuleList()
        for _ in range(operation_order.count("norm")):
            self.norms.append(build_norm_layer(norm_cfg, self.embed_dims)[1])

    def forward(self, query, key=None, value=None, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        norm_index = attn_index = ffn_index = 0
        identity = query
        if attn_masks is None:
            attn_masks = [None for _ in range(self.num_attn)]
        elif isinstance(attn_masks, torch.Tensor):
            attn_masks = [copy.deepcopy(attn_masks) for _ in range(self.num_attn)]
        for layer in self.operation_order:
            if layer == "self_attn":
                temp_key = temp_value = query
                query = self.attentions[attn_index](
                    query, temp_key, temp_value, identity if self.pre_norm else None,
                    query_pos=query_pos, key_pos=query_pos, attn_mask=attn_masks[attn_index],
                    key_padding_mask=query_key_padding_mask, **kwargs)
                attn_index += 1
                identity = query
            elif layer == "norm":
                query = self.norms[norm_index](query)
                norm_index += 1
            elif layer == "cross_attn":
                query = self.attentions[attn_index](
                    query, key, value, identity if self.pre_norm else None, query_pos=query_pos,
                    key_pos=key_pos, attn_mask=attn_masks[attn_index],
                    key_padding_mask=key_padding_mask, **kwargs)
                attn_index += 1
                identity = query
            elif layer == "ffn":
                query = self.ffns[ffn_index](query, identity if self.pre_norm else None)
                ffn_index += 1
        return query


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class TransformerLayerSequence(BaseModule):
    def __init__(self, transformerlayers=None, num_layers=None, init_cfg=None):
        super().__init__(init_cfg)
        if isinstance(transformerlayers, dict):
            transformerlayers = [copy.deepcopy(transformerlayers) for _ in range(num_layers)]
        self.num_layers = num_layers
        self.layers = ModuleList()
        for i in range(num_layers):
            self.layers.append(build_transformer_layer(dict(transformerlayers[i])))
        self.embed_dims = self.layers[0].embed_dims
        self.pre_norm = self.layers[0].pre_norm

    def forward(self, query, key, value, query_pos=None, key_pos=None, attn_masks=None,
                query_key_padding_mask=None, key_padding_mask=None, **kwargs):
        for layer in self.layers:
            query = layer(query, key, value, query_pos=query_pos, key_pos=key_pos,
                          attn_masks=attn_masks, query_key_padding_mask=query_key_padding_mask,
                          key_padding_mask=key_padding_mask, **kwargs)
        return query


# --------------------------------------------------------------------------- mmdet.models.utils.transformer
@TRANSFORMER_LAYER.register_module()
class DetrTransformerDecoderLayer(BaseTransformerLayer):
    def __init__(self, attn_cfgs, feedforward_channels, ffn_dropout=0.0, operation_order=None,
                 act_cfg=dict(type="ReLU", inplace=True), norm_cfg=dict(type="LN"), ffn_num_fcs=2,
                 **kwargs):
        super().__init__(attn_cfgs=attn_cfgs, feedforward_channels=feedforward_channels,
                         ffn_dropout=ffn_dropout, operation_order=operation_order, act_cfg=act_cfg,
                         norm_cfg=norm_cfg, ffn_num_fcs=ffn_num_fcs, **kwargs)
        assert len(operation_order) == 6


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class DetrTransformerEncoder(TransformerLayerSequence):
    def __init__(self, *args, post_norm_cfg=dict(type="LN"), **kwargs):
        super().__init__(*args, **kwargs)
        if post_norm_cfg is not None:
            self.post_norm = build_norm_layer(post_norm_cfg, self.embed_dims)[1] if self.pre_norm else None
        else:
            self.post_norm = None

    def forward(self, *args, **kwargs):
        x = super().forward(*args, **kwargs)
        if self.post_norm is not None:
            x = self.post_norm(x)
        return x


@TRANSFORMER_LAYER_SEQUENCE.register_module()
class DetrTransformerDecoder(TransformerLayerSequence):
    def __init__(self, *args, post_norm_cfg=dict(type="LN"), return_intermediate=False, **kwargs):
        super().__init__(*args, **kwargs)
        self.return_intermediate = return_intermediate
        self.post_norm = build_norm_layer(post_norm_cfg, self.embed_dims)[1] if post_norm_cfg else None


# --------------------------------------------------------------------------- assembling fake packages
class _Shim(types.ModuleType):
    """Module whose unknown attributes resolve to harmless no-op callables."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _noop


def _mod(name, **attrs):
    m = _Shim(name)
    m.__dict__.update(attrs)
    m.__path__ = []
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, m)
    return m


def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean, std, a, b)


def multi_apply(func, *args, **kwargs):
    from functools import partial
    pfunc = partial(func, **kwargs) if kwargs else func
    return tuple(map(list, zip(*map(pfunc, *args))))


_INSTALLED = False
BEV_POOL_IMPL = {"fn": None}


def _bev_pool_dispatch(feats, coords, B, D, H, W):
    return BEV_POOL_IMPL["fn"](feats, coords, B, D, H, W)


def install():
    """Put the fake mmcv/mmdet/mmdet3d packages and the reference namespace on sys.modules."""
    global _INSTALLED
    if _INSTALLED:
        return
    _INSTALLED = True
    if not hasattr(np, "int"):
        np.int = int  # reference uses np.int (occupancyformer.py:250)
    try:
        torch.backends.mha.set_fastpath_enabled(False)
    except Exception:
        pass

    NECKS, HEADS, BACKBONES, DETECTORS, LOSSES = (Registry(n) for n in
                                                  ("neck", "head", "backbone", "detector", "loss"))
    _mod("mmcv", deprecated_api_warning=_noop_decorator_factory, ConfigDict=ConfigDict)
    _mod("mmcv.utils", to_2tuple=lambda x: (x, x) if not isinstance(x, (tuple, list)) else tuple(x),
         Registry=Registry, build_from_cfg=build_from_cfg, ConfigDict=ConfigDict)
    _mod("mmcv.runner", BaseModule=BaseModule, ModuleList=ModuleList, Sequential=Sequential,
         force_fp32=_noop_decorator_factory, auto_fp16=_noop_decorator_factory)
    _mod("mmcv.cnn", build_norm_layer=build_norm_layer, build_conv_layer=build_conv_layer,
         ConvModule=ConvModule, Conv2d=nn.Conv2d, Conv3d=nn.Conv3d,
         build_activation_layer=build_activation_layer)
    _mod("mmcv.cnn.utils")
    _mod("mmcv.cnn.utils.weight_init", trunc_normal_=trunc_normal_)
    _mod("mmcv.cnn.bricks")
    _mod("mmcv.cnn.bricks.registry", ATTENTION=ATTENTION, POSITIONAL_ENCODING=POSITIONAL_ENCODING,
         TRANSFORMER_LAYER=TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE=TRANSFORMER_LAYER_SEQUENCE,
         FEEDFORWARD_NETWORK=FEEDFORWARD_NETWORK)
    _mod("mmcv.cnn.bricks.transformer", FFN=FFN, MultiheadAttention=MultiheadAttention,
         BaseTransformerLayer=BaseTransformerLayer, TransformerLayerSequence=TransformerLayerSequence,
         build_dropout=build_dropout, build_positional_encoding=build_positional_encoding,
         build_transformer_layer_sequence=build_transformer_layer_sequence,
         build_attention=build_attention, build_feedforward_network=build_feedforward_network,
         POSITIONAL_ENCODING=POSITIONAL_ENCODING, ATTENTION=ATTENTION,
         TRANSFORMER_LAYER=TRANSFORMER_LAYER, TRANSFORMER_LAYER_SEQUENCE=TRANSFORMER_LAYER_SEQUENCE)
    _mod("mmcv.ops")
    _mod("mmdet")
    _mod("mmdet.core", multi_apply=multi_apply, reduce_mean=lambda t: t)
    _mod("mmdet.core.anchor")
    _mod("mmdet.core.anchor.point_generator")
    _mod("mmdet.models", NECKS=NECKS, HEADS=HEADS, BACKBONES=BACKBONES, DETECTORS=DETECTORS)
    _mod("mmdet.models.builder", NECKS=NECKS, HEADS=HEADS, BACKBONES=BACKBONES, DETECTORS=DETECTORS,
         LOSSES=LOSSES, build_loss=lambda cfg: None)
    _mod("mmdet.models.backbones")
    _mod("mmdet.models.backbones.resnet", BasicBlock=BasicBlock)
    _mod("mmdet.models.utils")
    _mod("mmdet3d")
    _mod("mmdet3d.models")
    _mod("mmdet3d.models.builder", NECKS=NECKS, HEADS=HEADS, BACKBONES=BACKBONES, DETECTORS=DETECTORS)
    _mod("mmdet3d.ops")
    _mod("mmdet3d.ops.bev_pool", bev_pool=_bev_pool_dispatch)
    _mod("mmdet3d.ops.voxel_pooling")

    # reference namespace packages, registered WITHOUT running their __init__ (they drag in datasets)
    def ns(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m

    ns("projects", os.path.join(REFERENCE_ROOT, "projects"))
    ns("projects.mmdet3d_plugin", PLUGIN)
    utils = ns("projects.mmdet3d_plugin.utils", os.path.join(PLUGIN, "utils"))
    mu = load("projects.mmdet3d_plugin.utils.metric_util")
    utils.per_class_iu, utils.fast_hist_crop = mu.per_class_iu, mu.fast_hist_crop
    occ = os.path.join(PLUGIN, "occformer")
    ns("projects.mmdet3d_plugin.occformer", occ)
    for sub in ("backbones", "backbones/modules", "necks", "image2bev", "mask2former",
                "mask2former/base", "mask2former/positional_encodings"):
        ns("projects.mmdet3d_plugin.occformer." + sub.replace("/", "."), os.path.join(occ, sub))
    m = sys.modules["projects.mmdet3d_plugin.occformer.backbones.modules"]
    m.BottleNeckASPP = load("projects.mmdet3d_plugin.occformer.backbones.modules.aspp").BottleNeckASPP
    m.SwinBlock = load("projects.mmdet3d_plugin.occformer.backbones.modules.window_attention").SwinBlock

    # heads only inherit from these for type; forward is fully overridden (mask2former_nusc_occ.py:80)
    class AnchorFreeHead(BaseModule):
        pass

    class MaskFormerHead(AnchorFreeHead):
        pass

    base = "projects.mmdet3d_plugin.occformer.mask2former.base."
    _mod(base + "anchor_free_head", AnchorFreeHead=AnchorFreeHead)
    _mod(base + "maskformer_head", MaskFormerHead=MaskFormerHead)


def load(modname):
    """Import one reference file verbatim by dotted module name."""
    if modname in sys.modules and getattr(sys.modules[modname], "__file__", None):
        return sys.modules[modname]
    return importlib.import_module(modname)


def reference_available():
    return os.path.isdir(PLUGIN)
