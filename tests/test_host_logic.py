"""CPU tests of the product's host-side logic (no GPU, no kernels): parameter tables, table reductions, topology
bookkeeping and the host Philox stream — each checked against the oracle / the golden fixtures."""

import numpy as np
import pytest

import rcu_b200  # noqa: F401  (registers the package directory under this name)
from rcu_b200 import tables
from oracle import restate as R
from helpers import SWEEP, synth_metric_inputs, results_equal


def test_calibration_edges_are_exact_float32_ceilings():
    for n_bins in (1, 5, 10, 15, 32):
        e64 = np.linspace(0., 1. + 1e-8, n_bins + 1)
        e32 = tables.calibration_edges_f32(n_bins)
        assert e32.dtype == np.float32 and np.all(e32.astype(np.float64) >= e64)
        below = np.nextafter(e32[1:], np.float32(-1))
        assert np.all(below.astype(np.float64) < e64[1:])
    # the quirks SURVEY.md A5 lists: 0.5 -> bin 4, 1.0 -> bin 9
    p = np.array([0.5, 1.0, 0.7, 0.9], dtype=np.float32)
    assert list(np.digitize(p, np.linspace(0., 1. + 1e-8, 11)) - 1) == [4, 9, 6, 8]
    e32 = tables.calibration_edges_f32(10)
    assert list(np.searchsorted(e32, p, side='right') - 1) == [4, 9, 6, 8]


def test_float32_bin_rule_matches_digitize_on_adversarial_values():
    p, *_ = synth_metric_inputs(200000, 3)
    for n_bins in (10, 7, 32):
        ref = np.digitize(p, np.linspace(0., 1. + 1e-8, n_bins + 1)) - 1
        e32 = tables.calibration_edges_f32(n_bins)
        # the device rule: k = min(int(p*n), n-1); k += p >= e[k+1]; k -= p < e[k]
        k = np.minimum((p * np.float32(n_bins)).astype(np.int32), n_bins - 1)
        k = k + (p >= e32[k + 1])
        k = k - (p < e32[k])
        assert np.array_equal(k, ref)


def test_break_table_default_sweep():
    breaks, seg, order = tables.uncertainty_break_table()
    assert list(order) == list(range(11))
    assert len(breaks) >= 22 and np.all(np.diff(breaks.view(np.uint32).astype(np.int64)) > 0)
    assert seg[0] == 0 and seg[-1] == 0 and seg.max() == 11
    # classification by table == direct evaluation, on random values and on every float around each break
    p, *_ = synth_metric_inputs(400000, 5, with_break_neighbours=breaks)
    near = np.concatenate([(b.view(np.uint32).astype(np.int64) + np.arange(-300, 301)) for b in breaks]).astype(np.uint32).view(np.float32)
    p = np.concatenate([p, near])
    j = seg[np.searchsorted(breaks, p, side='right')]
    u = R.normalized_entropy(R.add_background_probability(p))
    assert np.array_equal(j, (u[:, None] > np.array(SWEEP)[None, :]).sum(1))


@pytest.mark.parametrize('ths', [(0.5,), (0.95, 0.05, 0.5), (-0.1, 0.5), (1.5,), (1e-6, 0.999), (0.3, 0.3)])
def test_break_table_arbitrary_thresholds(ths):
    breaks, seg, order = tables.uncertainty_break_table(ths)
    p, *_ = synth_metric_inputs(100000, 9, with_break_neighbours=breaks)
    j = seg[np.searchsorted(breaks, p, side='right')]
    u = R.normalized_entropy(R.add_background_probability(p))
    assert np.array_equal(j, (u[:, None] > np.sort(np.array(ths))[None, :]).sum(1))
    assert sorted(np.asarray(ths)[order]) == sorted(ths)


def test_break_table_rejects_untabulatable_threshold():
    with pytest.raises(ValueError):
        tables.uncertainty_break_table((1.0,))
    with pytest.raises(ValueError):
        tables.uncertainty_break_table(())


def test_threshold_breaks_f32_matches_numpy_comparison():
    u = np.random.default_rng(0).random(100000).astype(np.float32)
    ths = (0.05, 0.5, 0.95, 0.3)
    u[:4] = np.float32(ths)
    b, seg, order = tables.threshold_breaks_f32(ths)
    j = seg[np.searchsorted(b, u, side='right')]
    assert np.array_equal(j, sum((u > th).astype(np.int64) for th in ths))


@pytest.mark.parametrize('weighting', ['proportion', 'log_proportion', 'power_proportion', 'mean_proportion'])
def test_ece_from_tables_matches_golden(golden_metrics, weighting):
    p, target, mask = golden_metrics['p'], golden_metrics['target'], golden_metrics['mask']
    for with_mask in (False, True):
        count, positives, conf = R.calibration_tables(p, target, mask=mask if with_mask else None)
        bins = {}
        ece = tables.ece_from_tables(count, positives, conf, weighting, target.ndim, bins)
        pre = 'ece/mask%d/%s/' % (with_mask, weighting)
        results_equal(ece, golden_metrics[pre + 'ece'], 'ece')
        for k, v in bins.items():
            results_equal(v, golden_metrics[pre + k], k)
    with pytest.raises(ValueError):
        tables.ece_from_tables(count, positives, conf, 'nope')


def _joint_table(pred, target, unc, ths):
    j = (unc.reshape(-1)[:, None] > np.sort(np.array(ths))[None, :]).sum(1)
    t, d = target.reshape(-1).astype(bool), pred.reshape(-1).astype(bool)
    rows = (t & d, ~t & ~d, ~t & d, t & ~d)
    return np.array([[np.sum(r & (j == c)) for c in range(len(ths) + 1)] for r in rows], dtype=np.int64)


def test_correction_results_from_counts_match_golden(golden_metrics):
    pred, target, unc = golden_metrics['prediction'], golden_metrics['target'], golden_metrics['uncertainty']
    table = _joint_table(pred, target, unc, SWEEP)
    for k, th in enumerate(SWEEP):
        r = tables.correction_results(*tables.counts_at_threshold(table, k))
        keys = [key for key in golden_metrics.files if key.startswith('sweep/%s/' % th)]
        assert len(keys) == len(r) == 18
        for key in keys:
            results_equal(r[key.split('/')[-1]], golden_metrics[key], key)
        cols = tables.ue_table_columns(r)
        assert cols['ue'] == R.ue_table_row(r)['ue'] and cols['benefit'] == R.ue_table_row(r)['benefit']


def test_correction_results_degenerate(golden_metrics):
    unc = golden_metrics['uncertainty']
    z = np.zeros(unc.shape, dtype=np.uint8)
    table = _joint_table(z, z, unc, (0.5,))
    r = tables.correction_results(*tables.counts_at_threshold(table, 0))
    for key in [k for k in golden_metrics.files if k.startswith('degenerate_empty/0.5/')]:
        results_equal(r[key.split('/')[-1]], golden_metrics[key], key)


def test_error_ratios_edge_cases():
    assert tables.error_dice(0, 0, 0, 0, 0, 0) == 1. and tables.error_recall(0, 0, 0, 0) == 1. and tables.error_precision(0, 0, 0, 0) == 1.
    assert tables.dice_from_counts(0, 0, 0) == 1. and tables.accuracy_from_counts(0, 0, 0, 0) == 0
    assert tables.error_dice(2, 2, 1, 1, 1, 1) == R.error_dice(2, 2, 1, 1, 1, 1)


def test_unit_layout_matches_oracle_topology():
    from rcu_b200 import model
    for kw in (dict(), dict(in_channels=3), dict(dropout=0.5, dropout_center=4), dict(dropout=0.5, dropout_center=2),
               dict(dropout=None), dict(depth=3, start_filters=64)):
        cfg = R.UNetConfig(**kw)
        units, upconvs = model.unit_layout(cfg.in_channels, cfg.depth, cfg.start_filters, cfg.dropout, cfg.dropout_center)
        assert units == R.conv_sites(cfg)
        assert len(upconvs) == cfg.depth and upconvs[0][1] == cfg.start_filters * 2 ** cfg.depth


def test_host_philox_stream_matches_oracle():
    from rcu_b200 import metrics
    cfg = R.UNetConfig()
    sites = [c for _, c in R.dropout_sites(cfg)]
    scale = metrics.philox_keep_scale_host(20, cfg.dropout, sites, 100, 3, 2, 2)
    assert scale.shape == (2, 3, 2976)
    for t in range(2):
        keep = np.concatenate([m.numpy() for m in R.philox_keep_masks(cfg, 20, 2 + t, 100, 3)], axis=1).astype(bool)
        assert np.array_equal(scale[t] > 0, keep)
    assert set(np.unique(scale)) <= {np.float32(0), np.float32(1 / (1 - np.float32(0.05)))}
    # batching independence: a later slice window reproduces the same rows
    again = metrics.philox_keep_scale_host(20, cfg.dropout, sites, 101, 2, 2, 2)
    assert np.array_equal(again, scale[:, 1:])
    # 64-bit seeds use both key words
    a = metrics.philox_keep_scale_host(20 + (1 << 32), cfg.dropout, sites, 0, 1, 0, 1)
    b = metrics.philox_keep_scale_host(20, cfg.dropout, sites, 0, 1, 0, 1)
    assert not np.array_equal(a, b)
    with pytest.raises(ValueError):
        metrics.philox_keep_scale_host(20, 1.5, sites, 0, 1, 0, 1)


def test_synthetic_weights_have_the_reference_state_dict_layout():
    """rcu_b200.synth (bench / smoke / profiling weights) must stay loadable wherever a reference state dict is: same keys,
    shapes and dtypes as the oracle's restatement of the reference constructor (pinned to the reference in
    tests/test_oracle_golden*.py), and a usable logit spread."""
    import torch
    from rcu_b200 import synth
    for kw in (dict(in_channels=4), dict(in_channels=3), dict(in_channels=5), dict(in_channels=4, sigma_out=True)):
        cfg = R.UNetConfig(**kw)
        ref = R.init_state_dict(cfg, 20)
        sd = synth.random_unet_state_dict(seed=20, **kw)
        assert list(sd.keys()) != [] and set(sd) == set(ref)
        for k, v in ref.items():
            assert tuple(sd[k].shape) == tuple(v.shape) and sd[k].dtype == v.dtype, k
        again = synth.random_unet_state_dict(seed=20, **kw)
        assert all(torch.equal(sd[k], again[k]) for k in sd)
    psd = synth.random_postnet_state_dict()
    pref = R.postnet_init_state_dict(32, 2, 3, 21)
    assert set(psd) == set(pref) and all(tuple(psd[k].shape) == tuple(pref[k].shape) for k in pref)
    x = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(0))
    logits = R.unet_forward(synth.random_unet_state_dict(in_channels=4, seed=20), x, R.UNetConfig(in_channels=4))
    assert torch.isfinite(logits).all() and logits.std() > 0.05


def test_engine_cache_follows_the_weights_and_dies_with_the_module(monkeypatch):
    """steps.engine_for caches the converted engine per reference module; the entry must be rebuilt when the module's weights
    change (context.load_from_checkpoint / load_state_dict into the same object between two evaluations) and must not keep
    the module — or the engine and its device buffers — alive."""
    import gc
    import torch
    from rcu_b200 import steps
    built = []

    class FakeEngine:
        pass

    def fake_from_reference(cls, module, device=None, **kw):
        built.append(FakeEngine())
        return built[-1]
    monkeypatch.setattr(steps.B200UNet, 'from_reference', classmethod(fake_from_reference))
    module = torch.nn.Sequential(torch.nn.Conv2d(2, 2, 3), torch.nn.BatchNorm2d(2))
    e1 = steps.engine_for(module)
    assert steps.engine_for(module) is e1 and len(built) == 1
    module.load_state_dict({k: v + 1 for k, v in module.state_dict().items()})          # same object, new weights
    e2 = steps.engine_for(module)
    assert e2 is not e1 and len(built) == 2 and steps.engine_for(module) is e2
    with torch.no_grad():
        module[1].running_mean.add_(0.5)                                                 # a buffer written in place counts too
    assert steps.engine_for(module) is not e2 and len(built) == 3
    n_before = len(steps._ENGINES)
    del module
    gc.collect()
    assert len(steps._ENGINES) == n_before - 1
