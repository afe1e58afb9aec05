"""Host-side logic of the N>1 path on CPU: sample sharding + the single packed metric all-gather (gloo, world 2)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import port
from occformer_b200 import dist_eval

K = 17


def _data(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    target = torch.randint(0, K, (n, 10, 9, 4), generator=g)
    target[torch.rand(target.shape, generator=g) < 0.05] = 255
    pred = torch.where(torch.rand(target.shape, generator=g) < 0.6, target.clamp(max=K - 1), torch.randint(0, K, target.shape, generator=g))
    return pred, target


def test_counts_equal_reference_loops():
    pred, target = _data(3)
    assert torch.equal(dist_eval.ssc_counts(pred, target, K), port.ssc_counts_ref(pred, target, K))
    s = dist_eval.ssc_scores(dist_eval.ssc_counts(pred, target, K), K)
    assert 0.0 < s["iou_ssc_mean"] < 1.0 and 0.0 < s["iou"] <= 1.0


def test_shard_indices_cover_all_samples_once():
    for n, w in [(32, 8), (7, 2), (5, 4), (1, 2)]:
        seen = sum((dist_eval.shard_indices(n, r, w) for r in range(w)), [])
        assert seen == list(range(n))


def _worker(rank, world, port_no, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pred, target = _data(n)
    idx = dist_eval.shard_indices(n, rank, world)
    vec = dist_eval.ssc_counts(pred[idx], target[idx], K) if idx else torch.zeros(3 + 3 * K, dtype=torch.long)
    total = dist_eval.reduce_counts(vec)
    if rank == 0:
        torch.save(total, out)
    dist.destroy_process_group()


def test_sharded_eval_gloo_world2(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_no = s.getsockname()[1]; s.close()
    n, out = 5, str(tmp_path / "total.pt")
    mp.spawn(_worker, args=(2, port_no, n, out), nprocs=2, join=True)
    pred, target = _data(n)
    assert torch.equal(torch.load(out), dist_eval.ssc_counts(pred, target, K)), "sum over ranks != single-process counts"
