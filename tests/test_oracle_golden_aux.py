"""Pins the oracle's restatement of the single-forward methods that reuse the conv engine (SURVEY.md §8f rank 3-4:
sigma head / aleatoric step, features + PostNet, 5-channel auxiliary input, border mask, confidence preparations) to
tests/golden/aux_golden.npz, which tests/golden/make_golden_aux.py produced with the UNMODIFIED reference."""
import numpy as np
import pytest
import torch

from oracle import restate as R

torch.set_grad_enabled(False)


def aleatoric_state(golden_aux=None):
    cfg = R.UNetConfig(in_channels=4, sigma_out=True)
    sd0 = R.init_state_dict(cfg, 20)
    sd = R.randomize_statistics(sd0, 7)
    sd['conv_sigma.1.weight'] = sd['conv_sigma.1.weight'] * 6.0
    return cfg, sd0, sd


def auxfeat_state():
    cfg = R.UNetConfig(in_channels=4)
    sd = R.randomize_statistics(R.init_state_dict(cfg, 20), 7)
    psd0 = R.postnet_init_state_dict(32, 2, 3, 21)
    psd = R.randomize_statistics(psd0, 9)
    psd['conv_logits.weight'] = psd['conv_logits.weight'] * 4.0
    return cfg, sd, psd0, psd


def test_sigma_net_constructor_and_aleatoric_step(golden_aux):
    cfg, sd0, sd = aleatoric_state()
    assert np.float64(sum(v.double().sum().item() for v in sd0.values())) == golden_aux['aleatoric/param_sum']
    assert np.array_equal(sd0['conv_sigma.1.weight'].numpy(), golden_aux['aleatoric/sigma_head_weight'])
    assert len(R.dropout_sites(cfg)) == 20 and R.dropout_sites(cfg)[-1] == (R.SIGMA_PREFIX, 32)
    x = torch.from_numpy(golden_aux['aleatoric/input'])
    for is_log in (False, True):
        out = R.predict_aleatoric(sd, x, cfg, is_log_sigma=is_log)
        for k in ('logits', 'sigma', 'probabilities'):
            assert np.array_equal(out[k].numpy(), golden_aux['aleatoric/log%d/%s' % (is_log, k)]), (is_log, k)
    assert (golden_aux['aleatoric/log0/sigma'] >= 0).all() and golden_aux['aleatoric/log0/sigma'].std() > 0.05


def test_postnet_constructor_and_auxiliary_feature_step(golden_aux):
    cfg, sd, psd0, psd = auxfeat_state()
    assert np.float64(sum(v.double().sum().item() for v in psd0.values())) == golden_aux['auxfeat/postnet_param_sum']
    assert np.array_equal(psd0['conv_logits.weight'].numpy(), golden_aux['auxfeat/postnet_logits_weight'])
    out = R.predict_aux_feat(sd, cfg, psd, torch.from_numpy(golden_aux['auxfeat/input']))
    for k in ('segm_probabilities', 'probabilities', 'features'):
        assert np.array_equal(out[k].numpy(), golden_aux['auxfeat/' + k]), k
    p = golden_aux['auxfeat/probabilities']
    assert p.std() > 0.05   # the fixture is not a constant map


def test_auxiliary_segmentation_step(golden_aux):
    cfg = R.UNetConfig(in_channels=5)
    sd = R.randomize_statistics(R.init_state_dict(cfg, 20), 7)
    out = R.predict_aux_segm(sd, torch.from_numpy(golden_aux['auxsegm/input']), torch.from_numpy(golden_aux['auxsegm/labels']), cfg)
    for k in ('logits', 'probabilities', 'orig_prediction'):
        assert np.array_equal(out[k].numpy(), golden_aux['auxsegm/' + k]), k


@pytest.mark.parametrize('d_in,d_out', [(1, 1), (2, 1), (1, 2), (3, 3), (0, 1)])
def test_border_mask(golden_aux, d_in, d_out):
    dist, mask = R.boarder_mask(golden_aux['border/label'], d_in, d_out)
    assert np.array_equal(mask, golden_aux['border/mask_%d_%d' % (d_in, d_out)])
    if (d_in, d_out) == (1, 1):
        assert np.array_equal(dist, golden_aux['border/dist'])
        assert 0 < mask.sum() < mask.size
        _, m2 = R.boarder_mask(golden_aux['border/label'][3], 1, 1)
        assert np.array_equal(m2, golden_aux['border/mask2d_1_1'])


def test_confidence_preparations(golden_aux):
    u, pred = golden_aux['prep/uncertainty'], golden_aux['prep/prediction']
    resc = R.rescale_uncertainties(u, float(u.min()), float(u.max()))
    assert resc.dtype == golden_aux['prep/rescaled'].dtype and np.array_equal(resc, golden_aux['prep/rescaled'])
    fg = R.uncertainty_to_foreground_probabilities(resc.copy(), pred)
    assert np.array_equal(fg, golden_aux['prep/foreground'])
    with pytest.raises(ValueError):
        R.uncertainty_to_foreground_probabilities(u, pred)          # not rescaled: values > 1
    with pytest.raises(ValueError):
        R.uncertainty_to_foreground_probabilities(resc, pred[:2])


def test_residual_net_restatement_matches_the_unmodified_reference():
    """residual=True (ConvResidualBlock, common/model/unet.py:42-60): parameter creation order (so a seeded init is reproduced
    bit for bit), state_dict key order, the deterministic forward, `UNet.features` and a stochastic forward with injected
    Dropout2d masks — against tests/golden/residual_golden.npz, written by the unmodified reference UNet."""
    import os
    import torch
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'residual_golden.npz'))
    cfg = R.UNetConfig(in_channels=4, residual=True)
    sd = R.init_state_dict(cfg, 20)
    assert [str(k) for k in g['keys']] == list(sd.keys())
    assert np.float64(sum(v.double().sum().item() for v in sd.values())) == g['param_sum']
    assert np.float64(sum(v.double().abs().sum().item() for v in sd.values())) == g['param_abs_sum']
    assert np.array_equal(sd['down_convs.0.block.residual.weight'].numpy(), g['residual0_weight'])
    assert np.array_equal(sd['up_convs.3.block.residual.bias'].numpy(), g['residual_last_bias'])
    sd = R.randomize_statistics(sd, 7)
    x = torch.from_numpy(g['input'])
    with torch.no_grad():
        out = R.unet_forward(sd, x, cfg, return_all=True)
        mc = R.unet_forward(sd, x, cfg, R.philox_keep_masks(cfg, 20, 0, 0, x.shape[0]))
    assert np.allclose(out['logits'].numpy(), g['logits'], rtol=0, atol=2e-5)
    assert np.allclose(out['features'].numpy(), g['features'], rtol=0, atol=2e-5)
    assert np.allclose(mc.numpy(), g['mc_logits'], rtol=0, atol=2e-5)
    assert np.abs(g['mc_logits'] - g['logits']).max() > 1e-2          # the masks do something
