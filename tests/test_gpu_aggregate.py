"""GPU parity tests of the fused softmax / sample-mean / entropy / mutual-information / variance kernel against
the oracle's torch restatement of MultiPredictionSummary.  fp32 tolerance: 1e-6 absolute (probabilities and
entropies are O(1); the device uses the same formulas with expf/logf within 1-2 ulp of torch's CPU kernels)."""
import numpy as np
import pytest
import torch

from rcu_b200 import distributed as D
from rcu_b200 import metrics, steps
from oracle import restate as R

pytestmark = pytest.mark.gpu
TOL = 1e-6


@pytest.mark.parametrize('t,n,h,w', [(5, 3, 48, 64), (20, 2, 240, 240), (10, 1, 192, 256), (1, 1, 16, 16), (2, 1, 2, 2)])
def test_summary_from_logits_and_from_probabilities(t, n, h, w):
    torch.manual_seed(t * 100 + n)
    logits = (torch.randn(t, n, h, w, 2) * 3)
    logits[0, 0, 0, 0] = torch.tensor([40.0, -40.0])   # saturated softmax -> p == 0 exercises where(p > 0, ...)
    probs = torch.softmax(logits.permute(0, 1, 4, 2, 3), 2)
    ref = R.summarize(probs, do_mi=t > 1, do_var=t > 1)
    lazy = steps.LazyMultiProbabilities(logits.cuda())
    out = steps.summarize(lazy, do_mi=t > 1, do_var=t > 1, emit_prediction=True)
    out_p = steps.summarize(probs.cuda(), do_mi=t > 1, do_var=t > 1)
    for k, v in ref.items():
        assert out[k].shape == v.shape and out[k].dtype == torch.float32 and out[k].is_cuda
        assert (out[k].cpu() - v).abs().max().item() <= TOL, k
        assert (out_p[k].cpu() - v).abs().max().item() <= TOL, k
    assert torch.isfinite(out['entropy']).all()
    pred = (ref['probabilities'][:, 1] > ref['probabilities'][:, 0]).to(torch.uint8)
    clear = (ref['probabilities'][:, 1] - ref['probabilities'][:, 0]).abs() > 1e-5
    assert torch.equal(out['prediction'].cpu()[clear], pred[clear])
    assert (lazy.materialize().cpu() - probs).abs().max().item() <= TOL
    assert tuple(lazy.shape) == tuple(probs.shape)


def test_lazy_multi_probabilities_behaves_like_the_tensor():
    torch.manual_seed(0)
    logits = torch.randn(4, 2, 8, 8, 2).cuda()
    lazy = steps.LazyMultiProbabilities(logits)
    ref = torch.softmax(logits.permute(0, 1, 4, 2, 3), 2)
    assert torch.allclose(lazy.mean(dim=0), ref.mean(dim=0), atol=1e-6)           # attribute fall-through
    assert torch.allclose(torch.mean(lazy, dim=0), ref.mean(dim=0), atol=1e-6)    # __torch_function__


def test_partial_sums_then_finish_equals_one_pass():
    torch.manual_seed(1)
    logits = (torch.randn(7, 2, 32, 48, 2) * 2).cuda()
    one = steps.summarize(steps.LazyMultiProbabilities(logits), do_mi=True, do_var=True, emit_prediction=True)
    sums = D.aggregate_partial(logits[:3], True, True) + D.aggregate_partial(logits[3:], True, True)
    two = D.aggregate_finish(sums, 7, True, True, True)
    for k in ('probabilities', 'entropy', 'mutual_info'):
        assert (one[k] - two[k]).abs().max().item() <= 1e-6, k
    assert (one['variance'] - two['variance']).abs().max().item() <= 2e-6
    assert (one['prediction'] != two['prediction']).float().mean().item() < 1e-3


def test_philox_device_stream_equals_host_stream():
    host = metrics.philox_keep_scale_host(20, 0.05, [32, 32, 64, 512], 7, 5, 2, 3)
    dev = metrics.philox_keep_scale(20, 0.05, [32, 32, 64, 512], 7, 5, 2, 3).cpu().numpy()
    assert np.array_equal(host, dev)
    cfg = R.UNetConfig()
    keep = np.concatenate([m.numpy() for m in R.philox_keep_masks(cfg, 20, 4, 9, 2)], axis=1).astype(bool)
    dev = metrics.philox_keep_scale(20, cfg.dropout, [c for _, c in R.dropout_sites(cfg)], 9, 2, 4, 1).cpu().numpy()
    assert np.array_equal(dev[0] > 0, keep)
    assert abs(float((dev > 0).mean()) - 0.95) < 0.02


def test_bad_shapes_raise():
    with pytest.raises(ValueError):
        steps.summarize(torch.zeros(3, 2, 3, 4, 4).cuda())
    with pytest.raises(ValueError):
        steps.summarize(steps.LazyMultiProbabilities(torch.zeros(2, 1, 3, 5, 2).cuda()))  # odd H*W


@pytest.mark.parametrize('t,n,h,w', [(5, 3, 6, 10), (4, 2, 5, 6), (3, 40, 240, 240), (1, 1, 2, 2), (21, 2, 240, 240)])
def test_logit_difference_input_gives_bit_identical_outputs(t, n, h, w):
    """input_kind 3 (l0 - l1 per pixel, what the fused head writes in logit_diff mode) against input_kind 0 (the logit pairs):
    the two-class softmax is a function of the difference alone, so EVERY output is bit-identical — mean, entropy, MI,
    variance, argmax, foreground, the weight-scaling softmax riding in the launch and the materialised per-sample stack;
    small shapes (one pair per thread), hw not a multiple of four (8-byte loads) and the wide path (16-byte loads)."""
    torch.manual_seed(t * 7 + n)
    logits = (torch.randn(t + 1, n, h, w, 2) * 3).cuda()
    logits[1, 0, 0, 0] = torch.tensor([40.0, -40.0])
    logits[1, 0, 0, 1] = torch.tensor([-60.0, 60.0])
    diff = (logits[..., 0] - logits[..., 1]).contiguous()
    for mi_var in (False, True):
        if mi_var and t < 2:
            continue
        a = steps.summarize(steps.LazyMultiProbabilities(logits[1:]), do_mi=mi_var, do_var=mi_var, emit_prediction=True, emit_foreground=True,
                            ws_logits=logits[0])
        b = steps.summarize(steps.LazyMultiProbabilities(diff[1:], diff=True), do_mi=mi_var, do_var=mi_var, emit_prediction=True,
                            emit_foreground=True, ws_logits=diff[0])
        assert sorted(a) == sorted(b)
        for k in a:
            assert torch.equal(a[k], b[k]), k
        c = steps.summarize(steps.LazyMultiProbabilities(diff[1:], diff=True), do_mi=mi_var, do_var=mi_var)   # without the weight-scaling sample
        assert torch.equal(c['probabilities'], a['probabilities']) and torch.equal(c['entropy'], a['entropy'])
    assert torch.equal(steps.LazyMultiProbabilities(diff[1:], diff=True).materialize(), steps.LazyMultiProbabilities(logits[1:]).materialize())
    assert torch.equal(steps.softmax_planar(diff[0], diff=True), steps.softmax_planar(logits[0]))
    assert tuple(steps.LazyMultiProbabilities(diff[1:], diff=True).shape) == (t, n, 2, h, w)
