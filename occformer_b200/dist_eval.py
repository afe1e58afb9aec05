"""Sample-sharded evaluation: the ONLY cross-rank traffic of the inference path.

The forward pass shards by sample (every op is per sample); ranks exchange nothing until the end of
evaluation, when one packed int64 vector per rank is all-gathered and summed:

    [ lidarseg hist (K-1)x(K-1) | completion tp, fp, fn | semantic tp[K], fp[K], fn[K] ]     ((K-1)^2 + 3 + 3K int64, < 3 KB)

Reference semantics: SSCMetrics.get_score_completion / get_score_semantic_and_completion / compute
(projects/mmdet3d_plugin/utils/ssc_metric.py:104-168, 88-102), summed over ranks like the reference's
`dist.all_reduce(evaluation_semantic, SUM)` / `collect_results_cpu` (occformer/apis/test.py:195-212).  The counts
are computed from one confusion matrix (device kernel occ_ssc_counts / occ_lidarseg_hist) instead of the reference's K x 3
masked sums and numpy bincount -- same integers.
"""
import torch
import torch.distributed as dist


def ssc_counts(pred, target, num_classes, ignore=255):
    """pred, target: integer label volumes of identical shape (any leading batch dims).  -> int64 (3 + 3K,)
    CUDA tensors: one pass of the confusion-matrix kernel (occ_ssc_counts, csrc/eval_ops.cu).  CPU tensors (the host-side
    sharding / reduction logic is exercised on CPU under gloo): the same integers from torch.bincount."""
    K = num_classes
    if pred.is_cuda:
        from . import ops
        return ops.ssc_counts(pred, target, K, ignore)
    pred = pred.reshape(-1).long()
    target = target.reshape(-1).long()
    keep = target != ignore  # reference: predict[target==255] = 0; target[target==255] = 0, then mask = target != 255
    pred, target = pred[keep], target[keep]
    conf = torch.bincount(target * K + pred, minlength=K * K).view(K, K)  # conf[true, pred]
    tp = conf.diag()
    fp = conf.sum(0) - tp
    fn = conf.sum(1) - tp
    occ_t, occ_p = target > 0, pred > 0
    ctp = (occ_t & occ_p).sum()
    cfp = (~occ_t & occ_p).sum()
    cfn = (occ_t & ~occ_p).sum()
    return torch.cat([torch.stack([ctp, cfp, cfn]).long(), tp.long(), fp.long(), fn.long()])


def lidarseg_hist(point_scores, point_labels, num_classes, hist=None):
    """OccupancyFormer.simple_evaluation_semantic (occupancyformer.py:219-224,246-254): point_scores (n, K) from
    forward_lidarseg, point_labels (n,) integer in 0..K-1 (0 = unlabelled) -> (K-1, K-1) int64 hist[gt-1, pred-1] with
    pred = 1 + argmax(scores[:, 1:]); accumulated into `hist` when given."""
    K = num_classes
    if point_scores.is_cuda:
        from . import ops
        return ops.lidarseg_hist(point_scores, point_labels, K, hist)
    pred = point_scores[:, 1:].argmax(1) + 1
    gt = point_labels.long()
    keep = (gt >= 1) & (gt < K)
    h = torch.bincount((gt[keep] - 1) * (K - 1) + (pred[keep] - 1), minlength=(K - 1) ** 2).view(K - 1, K - 1)
    return h if hist is None else hist.add_(h)


def pack(ssc_vec, hist=None, num_classes=None):
    """The payload of the single collective: [ lidarseg hist (K-1)^2 | SC tp, fp, fn | SSC tp[K], fp[K], fn[K] ]"""
    if hist is None:
        hist = torch.zeros((num_classes - 1) ** 2, dtype=torch.int64, device=ssc_vec.device)
    return torch.cat([hist.reshape(-1).to(ssc_vec.device), ssc_vec])


def unpack(vec, num_classes):
    n = (num_classes - 1) ** 2
    return vec[:n].view(num_classes - 1, num_classes - 1), vec[n:]


def lidarseg_miou(hist):
    """per_class_iu (utils/metric_util.py:14-15): diag / (row + col - diag), mean over the K-1 classes"""
    h = hist.double()
    iu = h.diag() / (h.sum(1) + h.sum(0) - h.diag())
    return float(iu[~torch.isnan(iu)].mean()) if bool((~torch.isnan(iu)).any()) else float("nan")


def reduce_counts(vec, group=None):
    """One all_gather of the packed count vector; returns the sum over ranks (identity without a process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return vec.clone()
    parts = [torch.empty_like(vec) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, vec.contiguous(), group=group)
    return torch.stack(parts).sum(0)


def ssc_scores(vec, num_classes):
    """SSCMetrics.compute (ssc_metric.py:88-102) from the summed count vector."""
    v = vec.double()
    ctp, cfp, cfn = v[0], v[1], v[2]
    K = num_classes
    tp, fp, fn = v[3:3 + K], v[3 + K:3 + 2 * K], v[3 + 2 * K:3 + 3 * K]
    iou_ssc = tp / (tp + fp + fn + 1e-5)
    return {"precision": float(ctp / (ctp + cfp)), "recall": float(ctp / (ctp + cfn)),
            "iou": float(ctp / (ctp + cfp + cfn)), "iou_ssc": iou_ssc, "iou_ssc_mean": float(iou_ssc[1:].mean())}


def shard_indices(n_samples, rank, world):
    """Contiguous per-rank slices, like the reference's DistributedSampler for evaluation
    (projects/mmdet3d_plugin/datasets/samplers/distributed_sampler.py:35-39)."""
    per = (n_samples + world - 1) // world
    return list(range(rank * per, min((rank + 1) * per, n_samples)))
