"""Sample-sharded evaluation: the ONLY cross-rank traffic of the inference path.

The forward pass shards by sample (every op is per sample); ranks exchange nothing until the end of
evaluation, when one packed int64 vector per rank is all-gathered and summed:

    [ completion tp, fp, fn | semantic tp[K], fp[K], fn[K] ]            (3 + 3K int64, <= 3 KB)

Reference semantics: SSCMetrics.get_score_completion / get_score_semantic_and_completion / compute
(projects/mmdet3d_plugin/utils/ssc_metric.py:104-168, 88-102), summed over ranks like the reference's
`dist.all_reduce(evaluation_semantic, SUM)` / `collect_results_cpu` (occformer/apis/test.py:195-212).  The counts
are computed with one confusion-matrix bincount instead of the reference's K x 3 masked sums -- same integers.
"""
import torch
import torch.distributed as dist


def ssc_counts(pred, target, num_classes, ignore=255):
    """pred, target: integer label volumes of identical shape (any leading batch dims).  -> int64 (3 + 3K,)"""
    pred = pred.reshape(-1).long()
    target = target.reshape(-1).long()
    keep = target != ignore  # reference: predict[target==255] = 0; target[target==255] = 0, then mask = target != 255
    pred, target = pred[keep], target[keep]
    K = num_classes
    conf = torch.bincount(target * K + pred, minlength=K * K).view(K, K)  # conf[true, pred]
    tp = conf.diag()
    fp = conf.sum(0) - tp
    fn = conf.sum(1) - tp
    occ_t, occ_p = target > 0, pred > 0
    ctp = (occ_t & occ_p).sum()
    cfp = (~occ_t & occ_p).sum()
    cfn = (occ_t & ~occ_p).sum()
    return torch.cat([torch.stack([ctp, cfp, cfn]).long(), tp.long(), fp.long(), fn.long()])


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
