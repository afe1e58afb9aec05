"""Class-guided point sampling of the Mask2Former-3D heads -- SURVEY.md 8(a) rows A18 / A19 (training side).

    point_sample_3d, get_uncertainty, unravel_indices                          P/occformer/mask2former/base/mmdet_utils.py:21-90
    sample_valid_coords_with_frequencies / batch_...                           .../mmdet_utils.py:91-136
    get_nusc_lidarseg_point_coords                                             .../mmdet_utils.py:138-177
    get_uncertain_point_coords_3d_with_frequency                               .../mmdet_utils.py:179-246
    sampling weights (1 / class frequency, normalised, ** gamma)               P/occformer/mask2former/mask2former_occ.py:144-166

These functions are RNG driven (``torch.multinomial`` / ``torch.rand``) and belong to the training step; SURVEY.md 8(a)
keeps them on the host PyTorch path, outside the kernel scope and outside the parity gate (only distributional checks are
possible).  They are provided behind the reference's names and signatures so that training code written against the
plugin finds them here; ``grid_sample`` / ``multinomial`` / ``topk`` are torch library calls.
"""
import numpy as np
import torch
import torch.nn.functional as F

# voxel counts per SemanticKITTI class ('empty' first), P/utils/semkitti.py:3-26
SEMANTIC_KITTI_CLASS_FREQUENCIES = np.array([
    5.41773033e09, 1.57835390e07, 1.25136000e05, 1.18809000e05, 6.46799000e05, 8.21951000e05, 2.62978000e05, 2.83696000e05,
    2.04750000e05, 6.16887030e07, 4.50296100e06, 4.48836500e07, 2.26992300e06, 5.68402180e07, 1.57196520e07, 1.58442623e08,
    2.06162300e06, 3.69705220e07, 1.15198800e06, 3.34146000e05])


def class_sampling_weights(class_frequencies, gamma):
    """(1 / freq) / min(1 / freq), raised to gamma; gamma may be [lo, hi] for a per-call uniform draw (get_sampling_weights)."""
    w = 1.0 / np.asarray(class_frequencies, dtype=np.float64)
    w = w / w.min()
    if isinstance(gamma, (list, tuple)):
        gamma = np.random.uniform(low=gamma[0], high=gamma[1])
    return w ** gamma


def point_sample_3d(input, points, align_corners=False, **kwargs):
    """grid_sample with point coordinates in [0, 1]^3: input (N, C, D, H, W), points (N, P, 3) -> (N, C, P)
    (or points (N, a, b, c, 3) -> (N, C, a, b, c))."""
    flat = points.dim() == 3
    if flat:
        points = points[:, :, None, None, :]
    out = F.grid_sample(input, points * 2.0 - 1.0, align_corners=align_corners, **kwargs)
    return out[..., 0, 0] if flat else out


def get_uncertainty(mask_pred, labels):
    """-|logit| of the (class-specific, when there is more than one channel) prediction: (R, 1, ...)"""
    if mask_pred.shape[1] == 1:
        logits = mask_pred.clone()
    else:
        logits = mask_pred[torch.arange(mask_pred.shape[0], device=mask_pred.device), labels].unsqueeze(1)
    return -logits.abs()


def unravel_indices(indices, shape):
    """flat indices (*, N) -> coordinates (*, N, len(shape)) (row-major, like numpy.unravel_index)"""
    out = []
    for dim in reversed(tuple(shape)):
        out.append(indices % dim)
        indices = torch.div(indices, dim, rounding_mode="floor")
    return torch.stack(out[::-1], dim=-1)


def _voxel_weights(gt_labels, gt_masks, sample_weights):
    w = torch.as_tensor(sample_weights, device=gt_masks.device).float()
    return (w[gt_labels].view(-1, 1, 1, 1) * gt_masks).sum(dim=0).view(-1)


def _coords(point_indices, shape, like):
    coords = unravel_indices(point_indices, shape).float()
    norm = torch.tensor(tuple(shape)).type_as(like).view(1, 1, -1)
    return coords / (norm - 1).float()


@torch.no_grad()
def sample_valid_coords_with_frequencies(num_points, gt_labels, gt_masks, sample_weights=None):
    """num_points voxels of one sample drawn without replacement with probability ~ class weight of the voxel's label."""
    assert sample_weights is not None
    idx = torch.multinomial(_voxel_weights(gt_labels, gt_masks, sample_weights), num_samples=num_points, replacement=False)
    return idx, _coords(idx, gt_masks.shape[1:], gt_masks)


@torch.no_grad()
def batch_sample_valid_coords_with_frequencies(num_points, gt_labels_list, gt_masks_list, sample_weights=None):
    """The same draw, independently for every ground-truth instance of every sample: (sum num_gt, num_points) indices."""
    assert sample_weights is not None
    rows = []
    for gt_labels, gt_masks in zip(gt_labels_list, gt_masks_list):
        rows.append(_voxel_weights(gt_labels, gt_masks, sample_weights)[None].repeat(gt_labels.shape[0], 1))
    idx = torch.multinomial(torch.cat(rows, dim=0), num_samples=num_points, replacement=False)
    return idx, _coords(idx, gt_masks_list[-1].shape[1:], gt_masks_list[-1])


def _topk_uncertain(point_logits, labels, num_points, num_sampled, importance_sample_ratio):
    unc = get_uncertainty(point_logits.unsqueeze(1), labels)
    n_unc = int(importance_sample_ratio * num_points)
    idx = torch.topk(unc[:, 0, :], k=n_unc, dim=1)[1]
    idx = idx + num_sampled * torch.arange(point_logits.shape[0], dtype=torch.long, device=point_logits.device)[:, None]
    return idx, n_unc, num_points - n_unc


def get_nusc_lidarseg_point_coords(mask_pred, gt_lidarseg_list, labels, num_points, oversample_ratio, importance_sample_ratio,
                                   point_cloud_range, padding_mode="border", remove_noise_lidarseg=False):
    """nuScenes variant (A19): the sample's LiDAR points (normalised by the point-cloud range) topped up with uniform
    random points, over-sampled, then the most uncertain ones + fresh uniform ones.  -> (R, num_points, 3) in [0, 1]."""
    assert oversample_ratio >= 1 and 0 <= importance_sample_ratio <= 1
    R = mask_pred.shape[0]
    num_sampled = int(num_points * oversample_ratio)
    pcr = torch.tensor(point_cloud_range).type_as(mask_pred)
    per_row = []
    for i, pts in enumerate(gt_lidarseg_list):
        if remove_noise_lidarseg:
            pts = pts[pts[:, -1] > 0]
        c = (pts[:, :3] - pcr[:3]) / (pcr[3:] - pcr[:3])
        c = torch.cat((c, torch.rand((num_sampled - c.shape[0], 3), device=mask_pred.device)), dim=0)
        per_row.extend([c] * labels[i].shape[0])
    coords = torch.stack(per_row, dim=0)
    logits = point_sample_3d(mask_pred, coords[..., [2, 1, 0]], padding_mode=padding_mode).squeeze(1)
    idx, n_unc, n_rand = _topk_uncertain(logits, None, num_points, num_sampled, importance_sample_ratio)
    coords = coords.view(-1, 3)[idx.view(-1)].view(R, n_unc, 3)
    if n_rand > 0:
        coords = torch.cat((coords, torch.rand((R, n_rand, 3), device=mask_pred.device)), dim=1)
    return coords


def get_uncertain_point_coords_3d_with_frequency(mask_pred, labels, gt_labels_list, gt_masks_list, sample_weights, num_points,
                                                 oversample_ratio, importance_sample_ratio):
    """KITTI variant (A18): class-frequency guided over-sampling, the most uncertain points of it, plus uniformly drawn
    valid voxels.  -> (indices (R, num_points), coords (R, num_points, 3))."""
    assert oversample_ratio >= 1 and 0 <= importance_sample_ratio <= 1
    R = mask_pred.shape[0]
    num_sampled = int(num_points * oversample_ratio)
    pidx, pcoords = batch_sample_valid_coords_with_frequencies(num_sampled, gt_labels_list, gt_masks_list, sample_weights)
    if mask_pred.shape[-3:] == gt_masks_list[0].shape[1:]:
        logits = torch.gather(mask_pred.view(R, -1), dim=1, index=pidx)
    else:
        logits = point_sample_3d(mask_pred, pcoords[..., [2, 1, 0]], align_corners=True).squeeze(1)
    idx, n_unc, n_rand = _topk_uncertain(logits, labels, num_points, num_sampled, importance_sample_ratio)
    pidx = pidx.view(-1)[idx.view(-1)].view(R, n_unc)
    pcoords = pcoords.view(-1, 3)[idx.view(-1)].view(R, n_unc, 3)
    if n_rand > 0:
        ridx, rcoords = batch_sample_valid_coords_with_frequencies(n_rand, gt_labels_list, gt_masks_list,
                                                                   sample_weights=np.ones_like(np.asarray(sample_weights)))
        pidx, pcoords = torch.cat((pidx, ridx), dim=1), torch.cat((pcoords, rcoords), dim=1)
    return pidx, pcoords
