"""TEST INFRASTRUCTURE ONLY (oracle) -- CPU restatement of the OccFormer hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this file; the product package (occformer_b200/) never does.

Every function restates, in plain functional torch (fp32, CPU), one function of the
reference and cites the file:line it follows (paths relative to /root/reference,
P/ = projects/mmdet3d_plugin/, M/ = mmdetection3d/mmdet3d/).  Weights arrive as a
``state_dict`` carrying the reference's own keys.

Pinning: the reference ships no test or golden vector for this path ("parity
unpinned" by the reference itself); this port is pinned against the reference code
itself, imported verbatim under oracle/shim.py in the build container
(oracle/validate_port.py; tests/test_oracle_vs_reference.py) and through the
fixtures in tests/golden/ generated from the reference by oracle/gen_golden.py.
"""
import math

import torch
import torch.nn.functional as F

# =============================================================================
# LSS lift + voxel pooling
# =============================================================================


def voxel_index(geom, dx, bx):
    """P/occformer/image2bev/ViewTransformerLSSVoxel.py:84 -- fp32 subtract, fp32 divide,
    truncation toward zero (``.long()``).  ``bx - dx/2`` is evaluated in fp32 first."""
    return ((geom - (bx - dx / 2.0)) / dx).long()


def kept_mask(idx, nx):
    """ViewTransformerLSSVoxel.py:90-92 -- int64 index vs *float* nx, upper bound exclusive."""
    return ((idx[..., 0] >= 0) & (idx[..., 0] < nx[0]) & (idx[..., 1] >= 0) & (idx[..., 1] < nx[1])
            & (idx[..., 2] >= 0) & (idx[..., 2] < nx[2]))


def bev_pool(feats, coords, B, D, H, W):
    """M/ops/bev_pool/bev_pool.py:83-97 + src/bev_pool_cuda.cu:20-42 (the CUDA-only op).

    feats (n,C) fp32, coords (n,4) integer (x,y,z,b).  Returns (B,C,D,H,W) contiguous.
    Rank/sort/interval bookkeeping as in the reference; the kernel writes out[b,z,x,y,c]
    (bev_pool_cuda.cu:33-35) into a zero-initialised (B,D,H,W,C) buffer (bev_pool.cpp:38-41).
    Summation inside an interval is sequential in sorted order (bev_pool_cuda.cu:37-40); the
    reference's argsort is unstable so the order inside a voxel is unspecified -- we use a
    stable sort (a legal instance)."""
    assert feats.shape[0] == coords.shape[0]
    B_, D_, H_, W_ = (int(v) for v in (B, D, H, W))
    ranks = coords[:, 0] * (W * D * B) + coords[:, 1] * (D * B) + coords[:, 2] * B + coords[:, 3]
    indices = ranks.argsort(stable=True)
    feats, coords, ranks = feats[indices], coords[indices], ranks[indices]
    n, C = feats.shape
    out = feats.new_zeros((B_, D_, H_, W_, C))
    if n > 0:
        kept = torch.ones(n, dtype=torch.bool)
        kept[1:] = ranks[1:] != ranks[:-1]
        interval_starts = torch.where(kept)[0]
        seg = torch.cumsum(kept.long(), 0) - 1
        g = coords[interval_starts].long()
        sums = feats.new_zeros((interval_starts.numel(), C))
        sums.index_add_(0, seg, feats)  # sequential fp32 accumulation in sorted order on CPU
        out[g[:, 3], g[:, 2], g[:, 0], g[:, 1]] = sums
    return out.permute(0, 4, 1, 2, 3).contiguous()


def bev_pool_bookkeeping(coords, B, D, H, W):
    """Integer bookkeeping of bev_pool.py:86-93 / QuickCumsumCuda.forward :40-45:
    sorted unique ranks, interval_lengths (int64).  Exactness target for the CUDA path."""
    ranks = (coords[:, 0].long() * (int(W) * int(D) * int(B)) + coords[:, 1].long() * (int(D) * int(B))
             + coords[:, 2].long() * int(B) + coords[:, 3].long())
    uniq, counts = torch.unique(ranks, sorted=True, return_counts=True)
    return uniq, counts


def voxel_pooling(geom, volume, dx, bx, nx):
    """ViewTransformerLSSVoxel.voxel_pooling (ViewTransformerLSSVoxel.py:77-100).
    geom (B,N,D,fH,fW,3), volume (B,N,D,fH,fW,C).  Returns (B,C,X,Y,Z) (a permuted view, as
    in the reference :98) plus the integer bookkeeping (idx, kept)."""
    B, N, D, H, W, C = volume.shape
    Nprime = B * N * D * H * W
    x = volume.reshape(Nprime, C)
    idx = voxel_index(geom, dx, bx).view(Nprime, 3)
    batch_ix = torch.cat([torch.full([Nprime // B, 1], ix, dtype=torch.long) for ix in range(B)])
    gf = torch.cat((idx, batch_ix), 1)
    kept = kept_mask(gf, nx)
    final = bev_pool(x[kept], gf[kept], B, nx[2], nx[0], nx[1])
    return final.permute(0, 1, 3, 4, 2), gf, kept


def lift(depth_digit, img_feat, B, N):
    """ViewTransformerLSSVoxel.forward :110-115 -- depth softmax (x) context outer product.
    depth_digit (B*N,D,fH,fW), img_feat (B*N,C,fH,fW) -> volume (B,N,D,fH,fW,C), depth_prob."""
    depth_prob = depth_digit.softmax(dim=1)
    volume = depth_prob.unsqueeze(1) * img_feat.unsqueeze(2)
    _, C, D, H, W = volume.shape
    volume = volume.view(B, N, C, D, H, W).permute(0, 1, 3, 4, 5, 2)
    return volume, depth_prob


def gen_dx_bx(xbound, ybound, zbound):
    """P/occformer/image2bev/ViewTransformerLSSBEVDepth.py:21-25."""
    dx = torch.Tensor([row[2] for row in [xbound, ybound, zbound]])
    bx = torch.Tensor([row[0] + row[2] / 2.0 for row in [xbound, ybound, zbound]])
    nx = torch.Tensor([(row[1] - row[0]) / row[2] for row in [xbound, ybound, zbound]])
    return dx, bx, nx


def create_frustum(input_size, downsample, dbound):
    """ViewTransformerLSSBEVDepth.py:104-115."""
    ogfH, ogfW = input_size
    fH, fW = ogfH // downsample, ogfW // downsample
    ds = torch.arange(*dbound, dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = ds.shape[0]
    xs = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    ys = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((xs, ys, ds), -1)


def get_geometry(frustum, rots, trans, intrins, post_rots, post_trans, bda):
    """ViewTransformerLSSBEVDepth.get_geometry :117-150 (3x3 and KITTI 3x4 / 4x4 variants)."""
    B, N, _ = trans.shape
    points = frustum - post_trans.view(B, N, 1, 1, 1, 3)
    points = torch.inverse(post_rots).view(B, N, 1, 1, 1, 3, 3).matmul(points.unsqueeze(-1))
    points = torch.cat((points[:, :, :, :, :, :2] * points[:, :, :, :, :, 2:3],
                        points[:, :, :, :, :, 2:3]), 5)
    if intrins.shape[3] == 4:
        shift = intrins[:, :, :3, 3]
        points = points - shift.view(B, N, 1, 1, 1, 3, 1)
        intrins = intrins[:, :, :3, :3]
    combine = rots.matmul(torch.inverse(intrins))
    points = combine.view(B, N, 1, 1, 1, 3, 3).matmul(points).squeeze(-1)
    points = points + trans.view(B, N, 1, 1, 1, 3)
    if bda.shape[-1] == 4:
        points = torch.cat((points, torch.ones(*points.shape[:-1], 1).type_as(points)), dim=-1)
        points = bda.view(B, 1, 1, 1, 1, 4, 4).matmul(points.unsqueeze(-1)).squeeze(-1)
        points = points[..., :3]
    else:
        points = bda.view(B, 1, 1, 1, 1, 3, 3).matmul(points.unsqueeze(-1)).squeeze(-1)
    return points


# =============================================================================
# Dual-path voxel transformer encoder
# =============================================================================


def rel_position_index(ws=7):
    """WindowMSA.__init__ (P/occformer/backbones/modules/window_attention.py:57-61,109-113)."""
    seq1 = torch.arange(0, (2 * ws - 1) * ws, 2 * ws - 1)
    seq2 = torch.arange(0, ws, 1)
    coords = (seq1[:, None] + seq2[None, :]).reshape(1, -1)
    idx = coords + coords.T
    return idx.flip(1).contiguous()


def shift_attn_mask(H_pad, W_pad, ws=7, shift=3):
    """ShiftWindowMSA.forward mask construction (window_attention.py:186-208): region ids from
    slices (0,-ws),(-ws,-shift),(-shift,None); 0 / -100.0 additive mask, (nW, ws*ws, ws*ws)."""
    img_mask = torch.zeros((1, H_pad, W_pad, 1))
    slices = (slice(0, -ws), slice(-ws, -shift), slice(-shift, None))
    cnt = 0
    for h in slices:
        for w in slices:
            img_mask[:, h, w, :] = cnt
            cnt += 1
    mw = _window_partition(img_mask, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, float(-100.0)).masked_fill(am == 0, float(0.0))


def _window_partition(x, ws):
    """window_attention.py:260-274."""
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def _window_reverse(windows, H, W, ws):
    """window_attention.py:244-258."""
    B = int(windows.shape[0] / (H * W / ws / ws))
    x = windows.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def window_msa(x, sd, p, num_heads, mask=None, ws=7):
    """WindowMSA.forward (window_attention.py:69-107).  x (nW*B, 49, C)."""
    Bw, N, C = x.shape
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    qkv = qkv.reshape(Bw, N, 3, num_heads, C // num_heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q = q * (C // num_heads) ** -0.5
    attn = q @ k.transpose(-2, -1)
    table = sd[p + "relative_position_bias_table"]
    index = sd.get(p + "relative_position_index", rel_position_index(ws))
    bias = table[index.view(-1)].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if mask is not None:
        nW = mask.shape[0]
        attn = attn.view(Bw // nW, nW, num_heads, N, N) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, num_heads, N, N)
    attn = attn.softmax(dim=-1)
    x = (attn @ v).transpose(1, 2).reshape(Bw, N, C)
    return F.linear(x, sd[p + "proj.weight"], sd[p + "proj.bias"])


def shift_window_msa(query, hw_shape, sd, p, num_heads, shift_size, ws=7):
    """ShiftWindowMSA.forward (window_attention.py:168-242).  query (B', L, C), *already* LN'd:
    padding is applied after norm1 so pad tokens are exact zeros (Appendix D.6)."""
    B, L, C = query.shape
    H, W = hw_shape
    query = query.view(B, H, W, C)
    pad_r = (ws - W % ws) % ws
    pad_b = (ws - H % ws) % ws
    query = F.pad(query, (0, 0, 0, pad_r, 0, pad_b))
    H_pad, W_pad = query.shape[1], query.shape[2]
    if shift_size > 0:
        query = torch.roll(query, shifts=(-shift_size, -shift_size), dims=(1, 2))
        mask = shift_attn_mask(H_pad, W_pad, ws, shift_size)
    else:
        mask = None
    qw = _window_partition(query, ws).view(-1, ws * ws, C)
    aw = window_msa(qw, sd, p + "w_msa.", num_heads, mask, ws).view(-1, ws, ws, C)
    x = _window_reverse(aw, H_pad, W_pad, ws)
    if shift_size > 0:
        x = torch.roll(x, shifts=(shift_size, shift_size), dims=(1, 2))
    if pad_r > 0 or pad_b:
        x = x[:, :H, :W, :].contiguous()
    return x.view(B, H * W, C)


def swin_block(x, sd, p, num_heads, shift, ws=7):
    """SwinBlock.forward (window_attention.py:346-372).  x (B',C,H,W) NCHW in / out."""
    B, C, H, W = x.shape
    x = x.permute(0, 2, 3, 1).contiguous().view(B, -1, C)
    identity = x
    y = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    y = shift_window_msa(y, (H, W), sd, p + "attn.", num_heads, ws // 2 if shift else 0, ws)
    x = y + identity
    identity = x
    y = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    # mmcv FFN (hidden = C, GELU erf): layers.0.0 Linear, act, layers.1 Linear, + identity
    h = F.gelu(F.linear(y, sd[p + "ffn.layers.0.0.weight"], sd[p + "ffn.layers.0.0.bias"]))
    y = F.linear(h, sd[p + "ffn.layers.1.weight"], sd[p + "ffn.layers.1.bias"])
    x = identity + y
    return x.view(B, H, W, C).permute(0, 3, 1, 2).contiguous()


def _gn_relu(x, sd, p, groups, relu=True):
    x = F.group_norm(x, groups, sd[p + "weight"], sd[p + "bias"], 1e-5)
    return F.relu(x) if relu else x


def aspp_groups(C, norm_groups=32):
    """BottleNeckASPP.__init__ (P/occformer/backbones/modules/aspp.py:150-154)."""
    ch = C // 4
    return ch // 2 if ch <= norm_groups else norm_groups


def bottleneck_aspp(x, sd, p, norm_groups=32, dilations=(1, 6, 12, 18)):
    """BottleNeckASPP.forward + ASPP.forward (aspp.py:166-172, 107-122); eval => dropout = id."""
    C = x.shape[1]
    g_in = aspp_groups(C, norm_groups)
    identity = x
    y = F.conv2d(x, sd[p + "input_conv.0.weight"])
    y = _gn_relu(y, sd, p + "input_conv.1.", norm_groups)
    a = p + "aspp."
    inner = y
    x1 = _gn_relu(F.conv2d(y, sd[a + "aspp1.atrous_conv.weight"]), sd, a + "aspp1.bn.", g_in)
    branches = [x1]
    for i, d in zip((2, 3, 4), dilations[1:]):
        b = F.conv2d(y, sd[a + f"aspp{i}.atrous_conv.weight"], padding=d, dilation=d)
        branches.append(_gn_relu(b, sd, a + f"aspp{i}.bn.", g_in))
    x5 = F.adaptive_avg_pool2d(y, (1, 1))
    x5 = _gn_relu(F.conv2d(x5, sd[a + "global_avg_pool.1.weight"]), sd, a + "global_avg_pool.2.", g_in)
    x5 = F.interpolate(x5, size=y.shape[2:], mode="bilinear", align_corners=True)
    branches.append(x5)
    z = torch.cat(branches, dim=1)
    z = _gn_relu(F.conv2d(z, sd[a + "conv1.weight"]), sd, a + "bn1.", g_in)
    y = inner + z
    y = F.conv2d(y, sd[p + "output_conv.0.weight"])
    y = _gn_relu(y, sd, p + "output_conv.1.", norm_groups)
    return identity + y


def dualpath_block(x, sd, p, stride, shift, norm_groups=32):
    """DualpathTransformerBlock.forward (P/occformer/backbones/dualpath_block.py:65-82)."""
    identity = x
    w = sd[p + "input_conv.0.weight"]
    C = w.shape[0]
    num_heads = C // 32
    x = F.conv3d(x, w, None, stride=stride, padding=1)
    x = _gn_relu(x, sd, p + "input_conv.1.", norm_groups)
    x_bev = x.mean(dim=-1)
    B = x_bev.shape[0]
    Bz, Cc, X, Y, Z = x.shape
    xs = x.permute(0, 4, 1, 2, 3).reshape(B * Z, Cc, X, Y)  # 'b c x y z -> (b z) c x y'
    xs = torch.cat((x_bev, xs), dim=0)
    xs = swin_block(xs, sd, p + "bev_encoder.", num_heads, shift)
    x_bev, xs = xs[:B], xs[B:]
    x = xs.reshape(B, Z, Cc, X, Y).permute(0, 2, 3, 4, 1)  # '(b z) c x y -> b c x y z'
    x_bev = bottleneck_aspp(x_bev, sd, p + "aspp.", norm_groups)
    coeff = F.conv3d(x, sd[p + "combine_coeff.weight"], sd.get(p + "combine_coeff.bias")).sigmoid()
    x = x + coeff * x_bev.unsqueeze(-1)
    if stride > 1:
        idn = F.conv3d(identity, sd[p + "downsample.0.weight"], None, stride=stride)
        idn = _gn_relu(idn, sd, p + "downsample.1.", norm_groups, relu=False)
    else:
        idn = identity
    return x + idn


def occupancy_encoder(x, sd, block_numbers, block_strides, out_indices, prefix="", norm_groups=32):
    """OccupancyEncoder.forward (P/occformer/backbones/occnet.py:64-74); shift = global layer
    index odd (dualpath_block.py:30, occnet.py:49-60)."""
    res = []
    layer_index = 0
    for s, (nb, st) in enumerate(zip(block_numbers, block_strides)):
        for b in range(nb):
            x = dualpath_block(x, sd, f"{prefix}layers.{s}.{b}.", st if b == 0 else 1,
                               layer_index % 2 == 1, norm_groups)
            layer_index += 1
        if s in out_indices:
            res.append(x)
    return res


# =============================================================================
# Mask2Former-3D occupancy decoder head
# =============================================================================


def sine_pos3d(B, X, Y, Z, num_feats, temperature=10000, scale=2 * math.pi, eps=1e-6, offset=0.0):
    """SinePositionalEncoding3D.forward with an all-False mask, normalize=True
    (P/occformer/mask2former/positional_encodings/positional_encoding.py:58-108)."""
    not_mask = torch.ones((B, X, Y, Z), dtype=torch.int)
    x_embed = not_mask.cumsum(1, dtype=torch.float32)
    y_embed = not_mask.cumsum(2, dtype=torch.float32)
    z_embed = not_mask.cumsum(3, dtype=torch.float32)
    x_embed = (x_embed + offset) / (x_embed[:, -1:] + eps) * scale
    y_embed = (y_embed + offset) / (y_embed[:, :, -1:] + eps) * scale
    z_embed = (z_embed + offset) / (z_embed[:, :, :, -1:] + eps) * scale
    dim_t = torch.arange(num_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_feats)
    out = []
    for e in (x_embed, y_embed, z_embed):
        pe = e[..., None] / dim_t
        out.append(torch.stack((pe[..., 0::2].sin(), pe[..., 1::2].cos()), dim=5).view(B, X, Y, Z, -1))
    return torch.cat(out, dim=4).permute(0, 4, 1, 2, 3)


def _mha(query, key, value, sd, p, num_heads, attn_mask=None):
    """torch.nn.MultiheadAttention math path (packed in_proj), seq-first (L,B,E)."""
    L, B, E = query.shape
    S = key.shape[0]
    w, b = sd[p + "in_proj_weight"], sd[p + "in_proj_bias"]
    q = F.linear(query, w[:E], b[:E])
    k = F.linear(key, w[E:2 * E], b[E:2 * E])
    v = F.linear(value, w[2 * E:], b[2 * E:])
    hd = E // num_heads
    q = q.view(L, B * num_heads, hd).transpose(0, 1) * (hd ** -0.5)
    k = k.view(S, B * num_heads, hd).transpose(0, 1)
    v = v.view(S, B * num_heads, hd).transpose(0, 1)
    attn = torch.bmm(q, k.transpose(1, 2))
    if attn_mask is not None:  # bool, True = blocked
        attn = attn.masked_fill(attn_mask, float("-inf"))
    attn = attn.softmax(dim=-1)
    out = torch.bmm(attn, v).transpose(0, 1).contiguous().view(L, B, E)
    return F.linear(out, sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def decoder_layer(query, key, query_pos, key_pos, attn_mask, sd, p, num_heads):
    """mmcv BaseTransformerLayer with operation_order ('cross_attn','norm','self_attn','norm',
    'ffn','norm') (cfg occformer_nusc_r50_256x704.py:144-168; semantics SURVEY Appendix C):
    post-norm, V carries no positional encoding, self-attn uses key_pos = query_pos."""
    E = query.shape[-1]
    # cross attention: q = query+query_pos, k = key+key_pos, v = key
    out = _mha(query + query_pos, key + key_pos, key, sd, p + "attentions.0.attn.", num_heads, attn_mask)
    query = query + out
    query = F.layer_norm(query, (E,), sd[p + "norms.0.weight"], sd[p + "norms.0.bias"], 1e-5)
    out = _mha(query + query_pos, query + query_pos, query, sd, p + "attentions.1.attn.", num_heads)
    query = query + out
    query = F.layer_norm(query, (E,), sd[p + "norms.1.weight"], sd[p + "norms.1.bias"], 1e-5)
    h = F.relu(F.linear(query, sd[p + "ffns.0.layers.0.0.weight"], sd[p + "ffns.0.layers.0.0.bias"]))
    query = query + F.linear(h, sd[p + "ffns.0.layers.1.weight"], sd[p + "ffns.0.layers.1.bias"])
    return F.layer_norm(query, (E,), sd[p + "norms.2.weight"], sd[p + "norms.2.bias"], 1e-5)


def forward_head(decoder_out, mask_feature, target_size, sd, num_heads, prefix=""):
    """Mask2Former*OccHead.forward_head (P/occformer/mask2former/mask2former_nusc_occ.py:426-471)."""
    E = decoder_out.shape[-1]
    d = F.layer_norm(decoder_out, (E,), sd[prefix + "transformer_decoder.post_norm.weight"],
                     sd[prefix + "transformer_decoder.post_norm.bias"], 1e-5).transpose(0, 1)
    cls_pred = F.linear(d, sd[prefix + "cls_embed.weight"], sd[prefix + "cls_embed.bias"])
    m = d
    for i in (0, 2, 4):
        m = F.linear(m, sd[prefix + f"mask_embed.{i}.weight"], sd[prefix + f"mask_embed.{i}.bias"])
        if i < 4:
            m = F.relu(m)
    mask_pred = torch.einsum("bqc,bcxyz->bqxyz", m, mask_feature)
    pooled = F.adaptive_max_pool3d(mask_pred.float(), target_size).flatten(2)
    attn_mask = pooled.sigmoid() < 0.5
    attn_mask = attn_mask.unsqueeze(1).repeat((1, num_heads, 1, 1)).flatten(0, 1)
    return cls_pred, mask_pred, attn_mask, pooled


def head_forward(voxel_feats, sd, num_heads, num_layers, num_levels=3, prefix="", return_pooled=False):
    """Mask2FormerNuscOccHead.forward (mask2former_nusc_occ.py:589-689; KITTI twin
    mask2former_occ.py:569-671).  decoder_input_projs are Identity (feat == embed dims, :99-106)."""
    B = voxel_feats[0].shape[0]
    mask_features = voxel_feats[0]
    mems = voxel_feats[:0:-1]
    E = mask_features.shape[1]
    dec_in, dec_pos = [], []
    for i in range(num_levels):
        x = mems[i].flatten(2).permute(2, 0, 1)
        x = x + sd[prefix + "level_embed.weight"][i].view(1, 1, -1)
        X, Y, Z = mems[i].shape[-3:]
        pos = sine_pos3d(B, X, Y, Z, E / 3).flatten(2).permute(2, 0, 1)
        dec_in.append(x)
        dec_pos.append(pos)
    query_feat = sd[prefix + "query_feat.weight"].unsqueeze(1).repeat((1, B, 1))
    query_embed = sd[prefix + "query_embed.weight"].unsqueeze(1).repeat((1, B, 1))
    cls_list, mask_list, pooled_list = [], [], []
    cls_pred, mask_pred, attn_mask, pooled = forward_head(query_feat, mask_features,
                                                          mems[0].shape[-3:], sd, num_heads, prefix)
    cls_list.append(cls_pred), mask_list.append(mask_pred), pooled_list.append(pooled)
    for i in range(num_layers):
        lvl = i % num_levels
        attn_mask[torch.where(attn_mask.sum(-1) == attn_mask.shape[-1])] = False
        query_feat = decoder_layer(query_feat, dec_in[lvl], query_embed, dec_pos[lvl], attn_mask, sd,
                                   f"{prefix}transformer_decoder.layers.{i}.", num_heads)
        cls_pred, mask_pred, attn_mask, pooled = forward_head(
            query_feat, mask_features, mems[(i + 1) % num_levels].shape[-3:], sd, num_heads, prefix)
        cls_list.append(cls_pred), mask_list.append(mask_pred), pooled_list.append(pooled)
    if return_pooled:
        return cls_list, mask_list, pooled_list
    return cls_list, mask_list


def format_results(mask_cls, mask_pred):
    """mask2former_nusc_occ.py:691-696."""
    mask_cls = F.softmax(mask_cls, dim=-1)[..., :-1]
    return torch.einsum("bqc,bqxyz->bcxyz", mask_cls, mask_pred.sigmoid())


def forward_lidarseg(cls_preds, mask_preds, points, pc_range, padding_mode="border"):
    """mask2former_nusc_occ.py:505-542 (eval branch)."""
    pc_range = torch.tensor(pc_range).type_as(mask_preds)
    pc_min = pc_range[:3]
    ext = pc_range[3:] - pc_min
    voxel_preds = format_results(cls_preds, mask_preds)
    out = []
    for b, pts in enumerate(points):
        p = (pts[:, :3].float() - pc_min) / ext
        p = (p * 2) - 1
        p = p[..., [2, 1, 0]].view(1, 1, 1, -1, 3)
        s = F.grid_sample(voxel_preds[b:b + 1], p, mode="bilinear", padding_mode=padding_mode,
                          align_corners=True)
        out.append(s.squeeze().t().contiguous())
    return torch.softmax(torch.cat(out, dim=0), dim=1)


def head_simple_test(voxel_feats, sd, num_heads, num_layers, occ_size, num_levels=3, points=None,
                     pc_range=None, prefix=""):
    """Mask2FormerNuscOccHead.simple_test (mask2former_nusc_occ.py:698-745)."""
    cls_list, mask_list = head_forward(voxel_feats, sd, num_heads, num_layers, num_levels, prefix)
    mask_up = F.interpolate(mask_list[-1], size=tuple(occ_size), mode="trilinear", align_corners=True)
    res = {"output_voxels": [format_results(cls_list[-1], mask_up)], "output_points": None}
    if points is not None:
        res["output_points"] = forward_lidarseg(cls_list[-1], mask_list[-1], points, pc_range)
    return res


# =============================================================================
# MSDeformAttnPixelDecoder3D -- the neck between encoder and head (the first "next" row).
# ORACLE ONLY in this round: no CUDA implementation yet; pinned against the reference module under the shim
# (validate_port.check_neck) and by tests/golden/neck_small.npz.
# =============================================================================


def grid_priors_3d(shape, stride, offset=0.5):
    """MlvlPointGenerator.single_level_grid_priors, 3-D (P/utils/point_generator.py:111-135): cell centres in stride
    units, one row per voxel in (x slowest, z fastest) order, columns ordered (z, y, x) -- the order F.grid_sample wants
    for a (X, Y, Z) volume."""
    X, Y, Z = shape
    cx = (torch.arange(X, dtype=torch.float32) + offset) * stride
    cy = (torch.arange(Y, dtype=torch.float32) + offset) * stride
    cz = (torch.arange(Z, dtype=torch.float32) + offset) * stride
    gx, gy, gz = torch.meshgrid(cx, cy, cz, indexing="ij")
    return torch.stack([gz.reshape(-1), gy.reshape(-1), gx.reshape(-1)], dim=-1)


def ms_deform_attn_core_3d(value, shapes, loc, weights):
    """multi_scale_deformable_attn_pytorch (P/occformer/necks/multi_scale_deform_attn_3d.py:17-80).
    value (B, S, H, hd) with S = sum of level sizes; shapes [(X,Y,Z)] per level; loc (B, Nq, H, L, P, 3) in [0,1],
    last dim (z, y, x); weights (B, Nq, H, L, P).  Trilinear sampling (zeros outside, align_corners=False) of every
    level's value volume at the L*P points of every (query, head), weighted sum -> (B, Nq, H*hd)."""
    B, _, H, hd = value.shape
    Nq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    out = value.new_zeros(B * H, hd, Nq)
    start = 0
    for lvl, (X, Y, Z) in enumerate(shapes):
        n = X * Y * Z
        vol = value[:, start:start + n].permute(0, 2, 3, 1).reshape(B * H, hd, X, Y, Z)
        start += n
        grid = (2.0 * loc[:, :, :, lvl] - 1.0).permute(0, 2, 1, 3, 4).reshape(B * H, 1, Nq, P, 3)
        samp = F.grid_sample(vol, grid, mode="bilinear", padding_mode="zeros", align_corners=False)  # (BH,hd,1,Nq,P)
        w = weights[:, :, :, lvl].permute(0, 2, 1, 3).reshape(B * H, 1, Nq, P)
        out = out + (samp[:, :, 0] * w).sum(-1)
    return out.view(B, H * hd, Nq).transpose(1, 2).contiguous()


def ms_deform_attn_3d(query, query_pos, ref_points, shapes, sd, p, num_heads, num_points):
    """MultiScaleDeformableAttention3D.forward (multi_scale_deform_attn_3d.py:185-286) as called by BaseTransformerLayer
    'self_attn' with batch_first=False, post-norm: value = identity = the un-positioned query.  query (Nq, B, E)."""
    Nq, B, E = query.shape
    L = len(shapes)
    identity = query
    q = (query + query_pos).permute(1, 0, 2)
    v = F.linear(query.permute(1, 0, 2), sd[p + "value_proj.weight"], sd[p + "value_proj.bias"])
    v = v.view(B, Nq, num_heads, E // num_heads)
    off = F.linear(q, sd[p + "sampling_offsets.weight"], sd[p + "sampling_offsets.bias"])
    off = off.view(B, Nq, num_heads, L, num_points, 3)
    aw = F.linear(q, sd[p + "attention_weights.weight"], sd[p + "attention_weights.bias"])
    aw = aw.view(B, Nq, num_heads, L * num_points).softmax(-1).view(B, Nq, num_heads, L, num_points)
    norm = torch.tensor([[Z, Y, X] for (X, Y, Z) in shapes], dtype=query.dtype)  # offsets are in voxels of each level
    loc = ref_points[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    out = ms_deform_attn_core_3d(v, shapes, loc, aw)
    out = F.linear(out, sd[p + "output_proj.weight"], sd[p + "output_proj.bias"])
    return out.permute(1, 0, 2) + identity


def _conv_gn(x, sd, p, groups, padding=0, relu=False):
    """mmcv ConvModule(conv -> GN [-> ReLU]) with the reference's key names (conv.weight / conv.bias? / gn.*)."""
    x = F.conv3d(x, sd[p + "conv.weight"], sd.get(p + "conv.bias"), padding=padding)
    x = F.group_norm(x, groups, sd[p + "gn.weight"], sd[p + "gn.bias"], 1e-5)
    return F.relu(x) if relu else x


def ms_deform_pixel_decoder_3d(feats, sd, strides, num_heads, num_layers, num_levels=3, num_points=4,
                               norm_groups=32, prefix=""):
    """MSDeformAttnPixelDecoder3D.forward (P/occformer/necks/multiscale_deformattn_3d.py:143-248).
    feats: the 4 encoder outputs, high -> low resolution, (B, C_i, X_i, Y_i, Z_i).  Returns [mask_feature, multi-scale
    memories low-res last ...] in the reference's order (outs[::-1])."""
    nin = len(feats)
    B = feats[0].shape[0]
    E = sd[prefix + "level_encoding.weight"].shape[1]
    tokens, poss, refs, shapes = [], [], [], []
    for i in range(num_levels):  # encoder levels: coarsest first (:152-181)
        li = nin - i - 1
        f = feats[li]
        X, Y, Z = f.shape[-3:]
        proj = _conv_gn(f, sd, f"{prefix}input_convs.{i}.", norm_groups)
        pos = sine_pos3d(B, X, Y, Z, E // 3) + sd[prefix + "level_encoding.weight"][i].view(1, -1, 1, 1, 1)
        pts = grid_priors_3d((X, Y, Z), strides[li]) / (torch.tensor([[Z, Y, X]], dtype=torch.float32) * strides[li])
        tokens.append(proj.flatten(2).permute(2, 0, 1))
        poss.append(pos.flatten(2).permute(2, 0, 1))
        refs.append(pts)
        shapes.append((X, Y, Z))
    x = torch.cat(tokens, 0)
    qpos = torch.cat(poss, 0)
    ref = torch.cat(refs, 0)[None, :, None].repeat(B, 1, num_levels, 1)  # the same point for every level (:199-201)
    for l in range(num_layers):  # DetrTransformerEncoder of BaseTransformerLayer('self_attn','norm','ffn','norm')
        p = f"{prefix}encoder.layers.{l}."
        x = ms_deform_attn_3d(x, qpos, ref, shapes, sd, p + "attentions.0.", num_heads, num_points)
        x = F.layer_norm(x, (E,), sd[p + "norms.0.weight"], sd[p + "norms.0.bias"], 1e-5)
        h = F.relu(F.linear(x, sd[p + "ffns.0.layers.0.0.weight"], sd[p + "ffns.0.layers.0.0.bias"]))
        x = x + F.linear(h, sd[p + "ffns.0.layers.1.weight"], sd[p + "ffns.0.layers.1.bias"])
        x = F.layer_norm(x, (E,), sd[p + "norms.1.weight"], sd[p + "norms.1.bias"], 1e-5)
    mem = x.permute(1, 2, 0)
    outs, start = [], 0
    for (X, Y, Z) in shapes:
        n = X * Y * Z
        outs.append(mem[:, :, start:start + n].reshape(B, E, X, Y, Z))
        start += n
    for i in range(nin - num_levels - 1, -1, -1):  # FPN path for the levels that skipped the encoder (:228-246)
        cur = _conv_gn(feats[i], sd, f"{prefix}lateral_convs.{i}.", norm_groups)
        y = cur + F.interpolate(outs[-1], size=cur.shape[-3:], mode="trilinear", align_corners=False)
        outs.append(_conv_gn(y, sd, f"{prefix}output_convs.{i}.", norm_groups, padding=1, relu=True))
    outs[-1] = F.conv3d(outs[-1], sd[prefix + "mask_feature.weight"], sd[prefix + "mask_feature.bias"])
    return outs[::-1]


# the small neck configuration shared by validate_port.check_neck, gen_golden.gen_neck and the golden test
NECK_CASE = dict(in_channels=[32, 64, 128, 256], strides=[2, 4, 8, 16], E=96, layers=2, heads=4, levels=3, points=4,
                 ffn=192, sizes=[(16, 12, 8), (8, 6, 4), (4, 3, 2), (2, 2, 1)], wseed=21, xseed=23)


def neck_inputs(case, B=1):
    g = torch.Generator().manual_seed(case["xseed"])
    return [torch.randn(B, c, *sz, generator=g) for c, sz in zip(case["in_channels"], case["sizes"])]


# =============================================================================
# deterministic synthetic weights: shared with bench.py / smoke through occformer_b200.synth (pure data
# generators, no arithmetic of the path); re-exported here so tests keep one entry point
# =============================================================================
from occformer_b200.synth import make_block_state, make_encoder_state, make_head_state, make_neck_state  # noqa: E402,F401


# =============================================================================
# evaluation counts (the payload of the single metric all-gather)
# =============================================================================
def ssc_counts_ref(pred, target, num_classes):
    """SSCMetrics.get_score_completion + get_score_semantic_and_completion, literal loops
    (P/utils/ssc_metric.py:104-168) with mask = target != 255 and no nonempty mask."""
    pred, target = pred.clone().long(), target.clone().long()
    pred[target == 255] = 0
    valid = target != 255
    target[target == 255] = 0
    bs = pred.shape[0]
    p, t, m = pred.view(bs, -1), target.view(bs, -1), valid.view(bs, -1)
    ctp = cfp = cfn = 0
    tp = torch.zeros(num_classes, dtype=torch.long)
    fp, fn = tp.clone(), tp.clone()
    for i in range(bs):
        yt, yp = t[i][m[i]], p[i][m[i]]
        bt, bp = yt > 0, yp > 0
        ctp += int((bt & bp).sum()); cfp += int((~bt & bp).sum()); cfn += int((bt & ~bp).sum())
        for j in range(num_classes):
            tp[j] += ((yt == j) & (yp == j)).sum()
            fp[j] += ((yt != j) & (yp == j)).sum()
            fn[j] += ((yt == j) & (yp != j)).sum()
    return torch.cat([torch.tensor([ctp, cfp, cfn]), tp, fp, fn])
