"""Evaluation-tail kernels vs the oracle's literal loops / numpy-style bincount: exact integers."""
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu
K = 17


@pytest.mark.parametrize("shape", [(1, 10, 9, 4), (2, 50, 50, 8), (1, 33, 17, 5)])
def test_ssc_counts_kernel_exact(cuda, shape):
    from occformer_b200 import dist_eval
    g = torch.Generator().manual_seed(sum(shape))
    target = torch.randint(0, K, shape, generator=g)
    target[torch.rand(shape, generator=g) < 0.05] = 255
    pred = torch.where(torch.rand(shape, generator=g) < 0.6, target.clamp(max=K - 1), torch.randint(0, K, shape, generator=g))
    got = dist_eval.ssc_counts(pred.to(torch.uint8).to(cuda), target.to(torch.uint8).to(cuda), K)
    assert got.dtype == torch.int64 and got.is_cuda
    assert torch.equal(got.cpu(), port.ssc_counts_ref(pred, target, K))
    assert torch.equal(got.cpu(), dist_eval.ssc_counts(pred, target, K)), "device kernel != host bincount path"


def test_lidarseg_hist_kernel_exact(cuda):
    """fast_hist_crop semantics (utils/metric_util.py:8-23, occupancyformer.py:219-224,246-254)."""
    import numpy as np
    from occformer_b200 import dist_eval
    g = torch.Generator().manual_seed(1)
    n = 5000
    scores = torch.rand(n, K, generator=g)
    labels = torch.randint(0, K, (n,), generator=g)
    pred = (scores[:, 1:].argmax(1) + 1).numpy()
    gt = labels.numpy()
    nn = K  # np.max(unique_label) + 2 with unique_label = arange(16)
    k = (gt >= 0) & (gt < nn)
    hist = np.bincount(nn * gt[k].astype(int) + pred[k], minlength=nn ** 2)[:nn ** 2].reshape(nn, nn)[1:, 1:]
    got = dist_eval.lidarseg_hist(scores.to(cuda), labels.to(cuda), K)
    assert torch.equal(got.cpu(), torch.from_numpy(hist).long())
    again = dist_eval.lidarseg_hist(scores.to(cuda), labels.to(cuda), K, hist=got)  # accumulates
    assert torch.equal(again.cpu(), 2 * torch.from_numpy(hist).long())
    vec = dist_eval.pack(dist_eval.ssc_counts(torch.zeros(4, dtype=torch.uint8, device=cuda),
                                              torch.zeros(4, dtype=torch.uint8, device=cuda), K), got, K)
    h2, s2 = dist_eval.unpack(vec, K)
    assert vec.numel() == (K - 1) ** 2 + 3 + 3 * K and torch.equal(h2, got)
