"""Pins oracle/restate.py to the reference: every fixture under tests/golden/ was produced by the UNMODIFIED
reference code (tests/golden/make_golden.py), and the restatement must reproduce it bit for bit (same torch /
numpy on both boxes) — integers exactly, floats to the last bit unless stated."""
import numpy as np
import pytest
import torch

from oracle import restate as R
from helpers import GOLDEN_CONFIGS, SWEEP, results_equal

torch.set_grad_enabled(False)


@pytest.mark.parametrize('name', sorted(GOLDEN_CONFIGS))
def test_init_state_dict_matches_reference_constructor(golden_unet, name):
    cfg = R.UNetConfig(**GOLDEN_CONFIGS[name])
    sd = R.init_state_dict(cfg, 20)
    assert np.float64(sum(v.double().sum().item() for v in sd.values())) == golden_unet[name + '/param_sum']
    assert np.float64(sum(v.double().abs().sum().item() for v in sd.values())) == golden_unet[name + '/param_abs_sum']
    assert np.array_equal(sd['down_convs.0.block.block.0.conv2d_batch_relu.conv.weight'].numpy(), golden_unet[name + '/first_weight'])
    assert np.array_equal(sd['conv_cls.1.weight'].numpy(), golden_unet[name + '/head_weight'])


@pytest.mark.parametrize('name', sorted(GOLDEN_CONFIGS))
def test_unet_forward_and_steps_match_reference(golden_unet, name):
    cfg = R.UNetConfig(**GOLDEN_CONFIGS[name])
    sd = R.randomize_statistics(R.init_state_dict(cfg, 20), 7)
    x = torch.from_numpy(golden_unet[name + '/input'])
    det = R.predict_deterministic(sd, x, cfg)
    assert np.array_equal(det['logits'].numpy(), golden_unet[name + '/logits'])
    assert np.array_equal(det['probabilities'].numpy(), golden_unet[name + '/probabilities'])
    T = golden_unet[name + '/multi_probabilities'].shape[0]
    masks = [R.philox_keep_masks(cfg, 20, t, 0, x.shape[0]) for t in range(T)]
    mc = R.predict_mc(sd, x, cfg, T, masks)
    assert np.array_equal(mc['ws_probabilities'].numpy(), golden_unet[name + '/ws_probabilities'])
    assert np.array_equal(mc['multi_probabilities'].numpy(), golden_unet[name + '/multi_probabilities'])
    s = R.summarize(mc['multi_probabilities'], do_mi=True, do_var=True)
    for k in ('probabilities', 'entropy', 'mutual_info', 'variance'):
        assert np.array_equal(s[k].numpy(), golden_unet[name + '/summary_' + k]), k


def test_dropout_sites_counts():
    assert sum(c for _, c in R.dropout_sites(R.UNetConfig())) == 2976
    assert len(R.dropout_sites(R.UNetConfig())) == 19
    assert len(R.dropout_sites(R.UNetConfig(dropout=0.5, dropout_center=4))) == 9


def test_philox_known_answers():
    # Random123 known-answer vectors for philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, exp in kat:
        got = R.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(v) for v in got) == exp


def _to_eval(g):
    return g['p'], g['target'], g['mask'], g['prediction'], g['border'], g['uncertainty']


def test_eval_preparation(golden_metrics):
    p = golden_metrics['p']
    prob2 = R.add_background_probability(p)
    assert np.array_equal(prob2, golden_metrics['prob2'])
    u = R.normalized_entropy(prob2)
    assert u.dtype == np.float64 and np.array_equal(u, golden_metrics['uncertainty'])
    with pytest.raises(ValueError):
        R.add_background_probability(np.array([0.5, 1.5], dtype=np.float32))


@pytest.mark.parametrize('with_mask', [False, True])
@pytest.mark.parametrize('weighting', ['proportion', 'log_proportion', 'power_proportion', 'mean_proportion'])
def test_ece(golden_metrics, with_mask, weighting):
    p, target, mask, *_ = _to_eval(golden_metrics)
    ece, bins = R.ece_binary(golden_metrics['prob2'], target, mask=mask if with_mask else None, bin_weighting=weighting)
    pre = 'ece/mask%d/%s/' % (with_mask, weighting)
    results_equal(ece, golden_metrics[pre + 'ece'], 'ece')
    for k, v in bins.items():
        results_equal(v, golden_metrics[pre + k], k)


def test_ece_threshold_range(golden_metrics):
    ece, _ = R.ece_binary(golden_metrics['prob2'], golden_metrics['target'], threshold_range=(0.2, 0.8))
    results_equal(ece, golden_metrics['ece/range/ece'], 'ece')


def test_ece_rejects_multiclass():
    with pytest.raises(ValueError):
        R.calibration_tables(np.zeros((4, 3), dtype=np.float32), np.zeros(4, dtype=np.uint8))


def test_dice_confusion(golden_metrics):
    _, target, _, pred, _, _ = _to_eval(golden_metrics)
    tp, tn, fp, fn, n = R.confusion(pred, target)
    for k, v in (('tp', tp), ('tn', tn), ('fp', fp), ('fn', fn), ('n', n), ('dice', R.dice(pred, target))):
        results_equal(v, golden_metrics['dice_cm/' + k], k)


@pytest.mark.parametrize('th', SWEEP)
def test_sweep(golden_metrics, th):
    _, target, _, pred, border, unc = _to_eval(golden_metrics)
    r = R.uncertainty_and_correction(pred, target, unc, th)
    keys = [k for k in golden_metrics.files if k.startswith('sweep/%s/' % th)]
    assert len(keys) == 18
    for k in keys:
        results_equal(r[k.split('/')[-1]], golden_metrics[k], k)
    r = R.uncertainty_error_dice(pred, target, unc, th, 'ue_', border)
    for k in ('precision', 'recall', 'dice'):
        results_equal(r['ue_' + k], golden_metrics['uedice_border/%s/ue_%s' % (th, k)], k)
    r = R.uncertainty_error_dice(pred, target, unc, th)
    for k in ('precision', 'recall', 'dice'):
        results_equal(r[k], golden_metrics['uedice/%s/%s' % (th, k)], k)


def test_degenerate_empty_prediction(golden_metrics):
    _, _, _, _, _, unc = _to_eval(golden_metrics)
    z = np.zeros(unc.shape, dtype=np.uint8)
    r = R.uncertainty_and_correction(z, z, unc, 0.5)
    for k in [k for k in golden_metrics.files if k.startswith('degenerate_empty/0.5/')]:
        results_equal(r[k.split('/')[-1]], golden_metrics[k], k)


def test_restated_pymia_metrics_agree_with_an_independent_implementation():
    """pymia 0.2.1 (`ConfusionMatrix`, `DiceCoefficient`, `Accuracy`: numpyfunctions.py:128-151) is not installable here and is
    restated in oracle/ref_shim.py from its published semantics; the golden `dice_cm/*` entries therefore pin the
    restatement against itself.  scikit-learn IS installed: its confusion_matrix / f1_score / accuracy_score are an
    independent implementation of the same definitions (binary Dice == F1 of the positive class)."""
    from sklearn.metrics import accuracy_score, confusion_matrix, f1_score
    from oracle import ref_shim
    rng = np.random.default_rng(17)
    cases = [(rng.random(5000) < 0.3, rng.random(5000) < 0.25), (rng.random(777) < 0.9, rng.random(777) < 0.5),
             (np.zeros(64, bool), rng.random(64) < 0.5), (rng.random(64) < 0.5, np.zeros(64, bool)), (np.ones(10, bool), np.ones(10, bool))]
    for pred, target in cases:
        pred8, tgt8 = pred.astype(np.uint8), target.astype(np.uint8)
        tn, fp, fn, tp = confusion_matrix(tgt8, pred8, labels=[0, 1]).ravel()
        assert tuple(int(v) for v in R.confusion(pred8, tgt8)) == (tp, tn, fp, fn, pred.size)
        shim = ref_shim.ConfusionMatrix(pred8, tgt8)
        assert (int(shim.tp), int(shim.tn), int(shim.fp), int(shim.fn), shim.n) == (tp, tn, fp, fn, pred.size)
        assert np.isclose(R.dice(pred8, tgt8), f1_score(tgt8, pred8, zero_division=1.0), rtol=1e-15, atol=0)
        assert np.isclose(R.accuracy(pred8, tgt8), accuracy_score(tgt8, pred8), rtol=1e-15, atol=0)
    # the one convention sklearn cannot arbitrate: nothing predicted, nothing to find -> pymia's Dice is 1.0 (zero_division above)
    z = np.zeros(32, np.uint8)
    assert R.dice(z, z) == 1.0 and R.accuracy(z, z) == 1.0
