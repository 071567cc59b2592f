"""Parity properties at BASELINE.json's full size (one BraTS subject: 155 slices 4x240x240, T = 20 + the weight-scaling
pass; metrics over 8.93 M voxels).  The oracle needs minutes per slice-forward at this size, so the checks are the
size-independent properties the path offers: run-to-run bit determinism, invariance to how the call is split (chunks,
batches, sample ranges), linearity of the sample mean, probability / entropy range laws, count conservation and
additivity of the metric tables, monotonicity over the threshold sweep, invariance of the integer tables under a
permutation of the voxels."""
import numpy as np
import pytest
import torch

from rcu_b200 import distributed, metrics, model, steps, synth, tables

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
Z, H, W, T = 155, 240, 240, 20


@pytest.fixture(scope='module')
def subject():
    net = model.B200UNet(synth.random_unet_state_dict(in_channels=4, seed=20), in_channels=4, dropout=0.05, seed=20)
    x = torch.randn(Z, 4, H, W, generator=torch.Generator().manual_seed(7)).cuda()
    logits = net.forward_samples(x, T + 1, dropout_mode=1, det_first=True, seed=20, slice_index0=0)
    return net, x, logits


def test_forward_is_deterministic_and_split_invariant(subject):
    net, x, logits = subject
    assert logits.shape == (T + 1, Z, H, W, 2) and torch.isfinite(logits).all()
    again = net.forward_samples(x, T + 1, dropout_mode=1, det_first=True, seed=20, slice_index0=0)
    assert torch.equal(logits, again)                                  # bit-identical run to run
    del again
    # two calls (other chunk boundaries, other batch composition) == one call: the Philox stream is keyed by slice index
    tail = net.forward_samples(x[80:], T + 1, dropout_mode=1, det_first=True, seed=20, slice_index0=80)
    assert torch.equal(logits[:, 80:], tail)
    del tail
    # a sample range computed on its own (how ranks split the T samples) == the same samples of the folded call
    some = net.forward_samples(x[:16], 5, dropout_mode=1, seed=20, slice_index0=0, sample0=7)
    assert torch.equal(logits[8:13, :16], some)
    # the weight-scaling pass is the eval-mode forward, and MC samples really differ from it and from each other
    det = net.forward_samples(x[:16], 1, dropout_mode=0)
    assert torch.equal(det[0], logits[0, :16])
    assert not torch.equal(logits[1, :16], logits[0, :16]) and not torch.equal(logits[1, :16], logits[2, :16])
    # another seed is another stream
    other = net.forward_samples(x[:4], 2, dropout_mode=1, seed=21, slice_index0=0)
    assert not torch.equal(other, logits[1:3, :4])


def test_summary_laws_at_full_size(subject):
    _, _, logits = subject
    lazy = steps.LazyMultiProbabilities(logits[1:])
    out = steps.summarize(lazy, do_mi=True, do_var=True, emit_prediction=True, emit_foreground=True)
    p, ent, mi, var = out['probabilities'], out['entropy'], out['mutual_info'], out['variance']
    assert p.shape == (Z, 2, H, W) and ent.shape == mi.shape == var.shape == (Z, 1, H, W)
    assert (p.sum(1) - 1).abs().max().item() <= 1e-6 and p.min().item() >= 0 and p.max().item() <= 1
    assert ent.min().item() >= 0 and ent.max().item() <= np.log(2) + 1e-6          # th.entropy: nats, two classes
    assert mi.min().item() >= -1e-6 and (mi - ent).max().item() <= 1e-6           # 0 <= I <= H
    assert var.min().item() >= 0 and var.max().item() <= 0.25 * T / (T - 1) + 1e-6  # unbiased variance of values in [0, 1]
    assert torch.equal(out['foreground'], p[:, 1]) and torch.equal(out['prediction'].bool(), p[:, 1] > p[:, 0])
    # linearity of the sample mean: halves, and the partial-sum route the ranks use, give the same mean
    first = steps.summarize(steps.LazyMultiProbabilities(logits[1:11]))['probabilities']
    second = steps.summarize(steps.LazyMultiProbabilities(logits[11:]))['probabilities']
    assert ((first + second) * 0.5 - p).abs().max().item() <= 1e-6
    sums = distributed.aggregate_partial(logits[1:11], want_mi=True) + distributed.aggregate_partial(logits[11:], want_mi=True)
    fin = distributed.aggregate_finish(sums, T, has_mi=True, emit_prediction=True)
    assert (fin['probabilities'] - p).abs().max().item() <= 1e-6 and (fin['entropy'] - ent).abs().max().item() <= 1e-5
    assert (fin['mutual_info'] - mi).abs().max().item() <= 1e-5
    assert (fin['prediction'] != out['prediction']).sum().item() <= 8               # ties at p = 0.5 +- 1 ulp only
    # a second run of the fused pass is bit-identical
    assert torch.equal(steps.summarize(lazy)['probabilities'], p)


def test_metric_tables_conserve_and_add_up(subject):
    _, _, logits = subject
    out = steps.summarize(steps.LazyMultiProbabilities(logits[1:]), emit_prediction=True, emit_foreground=True)
    n = Z * H * W
    fg, pred = out['foreground'].reshape(-1), out['prediction'].reshape(-1)
    g = torch.Generator(device='cuda').manual_seed(3)
    # spread the probabilities over all bins (a random-init net sits in two of them) without leaving [0, 1]
    fg = (fg + torch.rand(n, device='cuda', generator=g)).remainder(1.0).contiguous()
    target = (torch.rand(n, device='cuda', generator=g) < fg).to(torch.uint8)
    mask = (torch.rand(n, device='cuda', generator=g) < 0.4).to(torch.uint8)
    count, positives, conf, ue, invalid, order = metrics.eval_fused(fg, pred, target, mask)
    assert invalid[0] == 0 and count[0, 10] == 0
    assert count[0].sum() == int(mask.sum().item())                                  # every masked voxel in exactly one bin
    assert positives[0].sum() == int((target.bool() & mask.bool()).sum().item()) and (positives[0] <= count[0]).all()
    lo, hi = np.arange(10) / 10, (np.arange(10) + 1) / 10
    mean_conf = conf[0, :10] / np.maximum(count[0, :10], 1)
    assert (count[0, :10] > 1000).all() and (mean_conf >= lo - 1e-6).all() and (mean_conf <= hi + 1e-6).all()
    assert np.isclose(conf[0, :10].sum(), fg.double()[mask.bool()].sum().item(), rtol=1e-9)
    assert ue[0].sum() == n                                                          # the U-E table is unmasked
    tp, tn, fp, fn = (int(v) for v in ue[0].sum(axis=1))
    assert (tp, tn, fp, fn) == tuple(int(v) for v in metrics.confusion_counts(pred, target)[0])
    assert tp == int(((pred == 1) & (target == 1)).sum().item()) and fn == int(((pred == 0) & (target == 1)).sum().item())
    # threshold sweep: U = {uncertainty > th} shrinks as th grows, so every *u count is non-increasing
    rows = [tables.counts_at_threshold(ue[0], k) for k in range(len(order))]
    for a, b in zip(rows, rows[1:]):
        assert a[:4] == b[:4] and all(x >= y for x, y in zip(a[4:], b[4:]))
    assert sum(rows[0][4:]) > sum(rows[-1][4:]) > 0
    # bit-identical on a second run; integer tables invariant under a permutation of the voxels, fp64 sums to 1e-12
    second = metrics.eval_fused(fg, pred, target, mask)
    for a, b in zip((count, positives, conf, ue), second[:4]):
        assert np.array_equal(a, b)
    perm = torch.randperm(n, device='cuda', generator=g)
    shuffled = metrics.eval_fused(fg[perm], pred[perm], target[perm], mask[perm])
    assert np.array_equal(count, shuffled[0]) and np.array_equal(positives, shuffled[1]) and np.array_equal(ue, shuffled[3])
    assert np.allclose(conf, shuffled[2], rtol=1e-12, atol=0)
    # additivity: the subject as 5 equal "subjects" of 31 slices in one launch sums to the whole
    parts = metrics.eval_fused(fg, pred, target, mask, n_subjects=5)
    assert np.array_equal(parts[0].sum(0), count[0]) and np.array_equal(parts[1].sum(0), positives[0])
    assert np.array_equal(parts[3].sum(0), ue[0]) and np.allclose(parts[2].sum(0), conf[0], rtol=1e-12, atol=0)
    # ECE from the tables equals the definition evaluated with torch on the device
    ece = tables.ece_from_tables(count[0, :10], positives[0, :10], conf[0, :10])
    m = mask.bool()
    b = torch.clamp((fg[m].double() * 10).floor().long(), max=9)
    acc = torch.zeros(10, dtype=torch.float64, device='cuda').index_add_(0, b, target[m].double())
    cnf = torch.zeros(10, dtype=torch.float64, device='cuda').index_add_(0, b, fg[m].double())
    cnt = torch.bincount(b, minlength=10).double()
    ref = ((acc / cnt - cnf / cnt).abs() * cnt / cnt.sum()).sum().item()
    assert abs(ece - ref) <= 1e-6        # bin edges at k/10 vs k*(1+1e-8)/10: a handful of voxels may sit in the neighbour bin
