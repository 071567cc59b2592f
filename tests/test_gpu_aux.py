"""GPU parity tests of the single-forward methods that reuse the conv engine against
tests/golden/aux_golden.npz (made by the UNMODIFIED reference, tests/golden/make_golden_aux.py) and the oracle:
aleatoric sigma head, auxiliary-feature PostNet (fused and through the materialised `features` tensor), 5-channel
auxiliary-segmentation input, border mask, confidence -> foreground-probability preparation.

Tolerances: the conv engine's (bf16 operands, fp32 accumulate; see test_gpu_unet.py); sigma / features are compared
relative to their own spread; byte masks and the float32 preparation arithmetic are bit-exact."""
import numpy as np
import pytest
import torch

from rcu_b200 import evaluation, metrics, model, steps
from oracle import restate as R
from test_oracle_golden_aux import aleatoric_state, auxfeat_state

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
P_MAX, P_MEAN = 1.2e-2, 2e-3


class Ctx:
    device = 'cuda'

    def __init__(self, model_):
        self.model = model_


class Batch:
    def __init__(self, **inputs):
        self.input, self.output, self.metrics = dict(inputs), {}, {}


def _prob_close(got, ref):
    got = got.cpu()
    ref = torch.as_tensor(ref)
    assert got.shape == ref.shape and got.dtype == torch.float32
    d = (got - ref).abs()
    assert d.max().item() <= P_MAX and d.mean().item() <= P_MEAN


def _rel_close(got, ref, tol_max=0.05, tol_mean=0.005):
    got, ref = got.cpu().float(), torch.as_tensor(ref).float()
    assert got.shape == ref.shape
    scale = ref.std().item()
    d = (got - ref).abs() / scale
    assert d.max().item() <= tol_max and d.mean().item() <= tol_mean, (d.max().item(), d.mean().item())


@pytest.mark.parametrize('is_log', [False, True])
def test_aleatoric_step_matches_reference_golden(golden_aux, is_log):
    cfg, _, sd = aleatoric_state()
    net = model.B200UNet(sd, in_channels=4, dropout=cfg.dropout)
    assert net.sigma_out and len(net.site_channels) == 20
    bc = Batch(images=torch.from_numpy(golden_aux['aleatoric/input']))
    steps.AleatoricPredictStep(is_log_sigma=is_log)(bc, None, Ctx(net))
    tag = 'aleatoric/log%d/' % is_log
    _prob_close(bc.output['probabilities'], golden_aux[tag + 'probabilities'])
    assert (bc.output['logits'].cpu() - torch.from_numpy(golden_aux[tag + 'logits'])).abs().max().item() <= 0.08
    ref_sigma = torch.from_numpy(golden_aux[tag + 'sigma'])
    if is_log:   # exp() amplifies: compare the raw head output
        _rel_close(bc.output['sigma'].log(), ref_sigma.log(), 0.08, 0.008)
    else:
        _rel_close(bc.output['sigma'], ref_sigma, 0.08, 0.008)
    assert bc.output['sigma'].shape == (2, 2, 48, 64) and (bc.output['sigma'] >= 0).all()
    # nn.Module protocol of a sigma_out net: model(x) -> (logits, sigma)   (unet.py:185-186)
    logits, sigma = net(bc.input['images'])
    assert torch.equal(logits, bc.output['logits']) and logits.shape == sigma.shape == (2, 2, 48, 64)


def test_sigma_branch_is_skipped_when_not_requested_and_mc_masks_cover_its_site():
    cfg, _, sd = aleatoric_state()
    net = model.B200UNet(sd, in_channels=4, dropout=cfg.dropout)
    x = torch.randn(2, 4, 32, 32)
    a = net.forward_outputs(x, 1, sigma=True)
    n_with = net.last_launch_count()
    b = net.forward_outputs(x, 1)
    assert net.last_launch_count() == n_with - 1 and torch.equal(a['logits'], b['logits'])
    # MC dropout through the sigma head: injected masks == the oracle with the same decisions
    masks = R.philox_keep_masks(cfg, 20, 0, 0, 2)
    assert len(masks) == 20
    out = net.forward_outputs(x, 1, dropout_mode=1, seed=20, sigma=True)
    ref = R.unet_forward(sd, x, cfg, masks, return_all=True)
    assert (out['logits'][0].permute(0, 3, 1, 2).cpu() - ref['logits']).abs().max().item() <= 0.08
    _rel_close(out['sigma'][0].permute(0, 3, 1, 2), ref['sigma'], 0.08, 0.008)
    det = R.unet_forward(sd, x, cfg, None, return_all=True)
    assert (ref['sigma'] - det['sigma']).abs().max().item() > 10 * (out['sigma'][0].permute(0, 3, 1, 2).cpu() - ref['sigma']).abs().max().item()


def test_auxiliary_feature_step_fused_and_materialised(golden_aux):
    cfg, sd, _, psd = auxfeat_state()
    seg = model.B200UNet(sd, in_channels=4, dropout=cfg.dropout, provide_features=True)
    post = model.B200PostNet(psd)
    assert post.nb_convs == 3 and post.in_channels == 32
    x = torch.from_numpy(golden_aux['auxfeat/input'])
    bc = Batch(images=x)
    steps.AuxiliaryFeatPredictStep(seg)(bc, None, Ctx(post))
    _prob_close(bc.output['segm_probabilities'], golden_aux['auxfeat/segm_probabilities'])
    _prob_close(bc.output['probabilities'], golden_aux['auxfeat/probabilities'])
    # the reference's own step body on the drop-in modules: model(features) with a materialised float32 tensor
    segm_logits = seg(x.cuda())
    assert seg.features.shape == (2, 32, 48, 64) and seg.features.dtype == torch.float32
    _rel_close(seg.features, golden_aux['auxfeat/features'], 0.08, 0.004)
    probs = torch.softmax(post(seg.features), 1)
    _prob_close(probs, golden_aux['auxfeat/probabilities'])
    assert (probs - bc.output['probabilities']).abs().max().item() <= 1e-5     # same bf16 features, same fp32 stack
    _prob_close(torch.softmax(segm_logits, 1), golden_aux['auxfeat/segm_probabilities'])
    # PostNet alone on the reference's exact features: fp32 CUDA-core arithmetic, tight agreement
    exact = torch.softmax(post(torch.from_numpy(golden_aux['auxfeat/features']).cuda()), 1)
    assert (exact.cpu() - torch.from_numpy(golden_aux['auxfeat/probabilities'])).abs().max().item() <= 2e-5
    with pytest.raises(ValueError):
        post(torch.zeros(1, 16, 8, 8))


def test_features_of_chunked_mc_batches_follow_the_sample_major_layout():
    cfg, sd, _, psd = auxfeat_state()
    seg = model.B200UNet(sd, in_channels=4, dropout=cfg.dropout, chunk_images=4)
    post = model.B200PostNet(psd)
    x = torch.randn(5, 4, 16, 32)
    out = seg.forward_outputs(x, 3, dropout_mode=1, det_first=True, seed=20, features=True, postnet=post)
    assert out['features'].shape == (3, 5, 32, 16, 32)
    big = model.B200UNet(sd, in_channels=4, dropout=cfg.dropout)
    ref = big.forward_outputs(x, 3, dropout_mode=1, det_first=True, seed=20, features=True, postnet=post)
    for k in ('logits', 'features', 'postnet_logits'):
        assert torch.equal(out[k], ref[k]), k
    again = post(out['features'].reshape(15, 32, 16, 32)).reshape(3, 5, 2, 16, 32)
    assert (again - out['postnet_logits'].permute(0, 1, 4, 2, 3)).abs().max().item() <= 1e-4


def test_auxiliary_segmentation_step_five_input_channels(golden_aux):
    cfg = R.UNetConfig(in_channels=5)
    sd = R.randomize_statistics(R.init_state_dict(cfg, 20), 7)
    net = model.B200UNet(sd, in_channels=5, dropout=cfg.dropout)
    bc = Batch(images=torch.from_numpy(golden_aux['auxsegm/input']), labels=torch.from_numpy(golden_aux['auxsegm/labels']))
    steps.AuxiliarySegmPredictStep()(bc, None, Ctx(net))
    _prob_close(bc.output['probabilities'], golden_aux['auxsegm/probabilities'])
    assert (bc.output['logits'].cpu() - torch.from_numpy(golden_aux['auxsegm/logits'])).abs().max().item() <= 0.08
    assert np.array_equal(bc.output['orig_prediction'].cpu().numpy(), golden_aux['auxsegm/orig_prediction'])
    assert bc.input['labels'].is_cuda and bc.input['labels'].dtype == torch.int64


@pytest.mark.parametrize('d_in,d_out', [(1, 1), (2, 1), (1, 2), (3, 3), (0, 1)])
def test_border_mask_equals_reference(golden_aux, d_in, d_out):
    label = golden_aux['border/label']
    got = evaluation.boarder_mask(label, d_in, d_out)
    assert got.dtype == bool and np.array_equal(got, golden_aux['border/mask_%d_%d' % (d_in, d_out)])
    dev = evaluation.boarder_mask(torch.from_numpy(label).cuda(), d_in, d_out)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy().astype(bool), got)
    if (d_in, d_out) == (1, 1):
        assert np.array_equal(evaluation.boarder_mask(label[3], 1, 1), golden_aux['border/mask2d_1_1'])


def test_border_mask_full_subject_against_oracle():
    rng = np.random.default_rng(3)
    zz, yy, xx = np.mgrid[0:155, 0:240, 0:240]
    label = (((zz - 70) / 40.0) ** 2 + ((yy - 120) / 60.0) ** 2 + ((xx - 100) / 50.0) ** 2) < 1
    label ^= rng.random(label.shape) < 0.001
    got = evaluation.boarder_mask(label.astype(np.uint8), 1, 1)
    _, ref = R.boarder_mask(label, 1, 1)
    assert np.array_equal(got, ref) and 0 < got.sum() < got.size // 10
    # the mask's consumer: UncertaintyErrorDiceNumpy(with_mask=True) evaluates outside the border (eval.py:163-165)
    with pytest.raises(NotImplementedError):
        metrics.border_mask(label, 9, 1)


def test_confidence_preparation_is_bit_exact(golden_aux):
    u, pred = golden_aux['prep/uncertainty'], golden_aux['prep/prediction']
    fg = evaluation.uncertainty_to_foreground_probabilities(u, pred, rescale=(float(u.min()), float(u.max())))
    assert fg.dtype == np.float32 and np.array_equal(fg, golden_aux['prep/foreground'])
    # already rescaled input, device tensors in -> device tensor out
    fg2 = evaluation.uncertainty_to_foreground_probabilities(torch.from_numpy(golden_aux['prep/rescaled']).cuda(),
                                                              torch.from_numpy(pred).cuda())
    assert fg2.is_cuda and np.array_equal(fg2.cpu().numpy(), golden_aux['prep/foreground'])
    # RescaleSubjectMinMax semantics (float32 min / max scalars, analysis.py:176)
    ref = R.uncertainty_to_foreground_probabilities(R.rescale_uncertainties(u, u.min(), u.max()), pred)
    got = evaluation.uncertainty_to_foreground_probabilities(u, pred, rescale='subject')
    assert np.array_equal(got, ref)
    mn, mx = metrics.minmax(np.concatenate([u.ravel(), np.array([-3.5, -0.0, 0.0], dtype=np.float32)]))
    assert mn == np.float32(-3.5) and mx == u.max()
    with pytest.raises(ValueError):
        evaluation.uncertainty_to_foreground_probabilities(u, pred)                  # values above 1
    with pytest.raises(ValueError):
        evaluation.uncertainty_to_foreground_probabilities(golden_aux['prep/rescaled'], pred * 2)
    with pytest.raises(ValueError):
        evaluation.uncertainty_to_foreground_probabilities(u, pred[:2])
    # downstream: the pseudo-probabilities feed the calibration tables like any foreground probability
    target = (np.random.default_rng(1).random(u.shape) < fg).astype(np.uint8)
    assert np.isclose(evaluation.ece_binary(fg, target), R.ece_binary(fg, target)[0], rtol=1e-12, atol=0)
