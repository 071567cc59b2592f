"""Generates the golden fixtures under tests/golden/ by running the UNMODIFIED reference (through
oracle/ref_shim.py) in the authoring container.  Run from the repo root:

    python tests/golden/make_golden.py

The fixtures pin oracle/restate.py (tests/test_oracle_golden.py) and are the reference side of the GPU parity
tests.  /root/reference does not exist on the GPU box, so nothing but this script reads it.
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, restate as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def synth_metric_inputs(n, seed):
    rng = np.random.default_rng(seed)
    p = rng.beta(0.3, 0.3, size=n).astype(np.float32)
    base = np.array([0.0, 1e-45, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1 - 2 ** -24], dtype=np.float32)
    adv = np.concatenate([base, np.nextafter(base, np.float32(2))[:-2], np.nextafter(base, np.float32(-1))[1:]])
    p[:len(adv)] = adv
    target = (rng.random(n) < p).astype(np.uint8)
    mask = rng.random(n) < 0.25
    prediction = (p > 0.5).astype(np.uint8)
    flip = rng.random(n) < 0.05
    prediction[flip] ^= 1
    border = rng.random(n) < 0.1
    return p, target, mask, prediction, border


def flatten_results(prefix, results, store):
    for k, v in results.items():
        store['{}/{}'.format(prefix, k)] = np.asarray(v)


def main():
    ref_shim.load()
    import common.model.unet as unet
    import common.trainloop.context as ctx
    import common.trainloop.steps as step
    import rechun.dl.customsteps as customsteps
    import common.evalutation.eval as ev
    import common.evalutation.numpyfunctions as np_fn
    import rechun.eval.analysis as analysis
    import torch.nn as nn
    torch.set_grad_enabled(False)
    warnings.simplefilter('ignore')

    # ------------------------------------------------------------------ U-Net / steps
    store = {}
    configs = {'brats': dict(in_channels=4), 'isic': dict(in_channels=3),
               'center': dict(in_channels=4, dropout=0.5, dropout_center=4)}
    for name, kw in configs.items():
        cfg = R.UNetConfig(**kw)
        torch.manual_seed(20)
        net = unet.UNet(cfg.nb_classes, cfg.in_channels, depth=cfg.depth, start_filters=cfg.start_filters,
                        dropout=cfg.dropout, dropout_center=cfg.dropout_center).eval()
        sd = net.state_dict()
        store[name + '/param_sum'] = np.float64(sum(v.double().sum().item() for v in sd.values()))
        store[name + '/param_abs_sum'] = np.float64(sum(v.double().abs().sum().item() for v in sd.values()))
        store[name + '/first_weight'] = sd['down_convs.0.block.block.0.conv2d_batch_relu.conv.weight'].numpy().copy()
        store[name + '/head_weight'] = sd['conv_cls.1.weight'].numpy().copy()
        net.load_state_dict(R.randomize_statistics(sd, 7))
        g = torch.Generator().manual_seed(1)
        x = torch.randn(2, cfg.in_channels, 48, 64, generator=g)
        store[name + '/input'] = x.numpy()
        c = ctx.TorchTestContext('cpu')
        c.model = net
        bc = ctx.BatchContext({'images': x.clone()}, 0)
        step.SegmentationPredictStep(do_probs=True)(bc, None, c)
        store[name + '/logits'] = bc.output['logits'].numpy()
        store[name + '/probabilities'] = bc.output['probabilities'].numpy()
        # MC with injected Philox keep masks: forward hooks replace the Dropout2d output (reference runs unmodified)
        T = 3
        drops = [m for m in net.modules() if isinstance(m, nn.Dropout2d)]
        state = {'t': -1}  # McPredictStep: call 0 is the deterministic pass (dropout in eval mode -> hook inactive)
        masks = [R.philox_keep_masks(cfg, 20, t, 0, x.shape[0]) for t in range(T)]
        calls = {'n': 0}

        def make_hook(site):
            def hook(mod, inp, out):
                if not mod.training:
                    return None
                t = calls['n'] // len(drops)
                calls['n'] += 1
                keep = masks[t][site].float()
                return inp[0] * (keep / (1 - mod.p))[:, :, None, None]
            return hook
        handles = [d.register_forward_hook(make_hook(i)) for i, d in enumerate(drops)]
        bc = ctx.BatchContext({'images': x.clone()}, 0)
        customsteps.McPredictStep(T)(bc, None, c)
        assert calls['n'] == T * len(drops)
        multi = bc.output['multi_probabilities'].clone()
        store[name + '/ws_probabilities'] = bc.output['ws_probabilities'].numpy()
        store[name + '/multi_probabilities'] = multi.numpy()
        customsteps.MultiPredictionSummary(do_mi=True, do_var=True)(bc, None, c)
        for k in ('probabilities', 'entropy', 'mutual_info', 'variance'):
            store[name + '/summary_' + k] = bc.output[k].numpy()
        for h in handles:
            h.remove()
    np.savez_compressed(os.path.join(OUT, 'unet_golden.npz'), **store)

    # ------------------------------------------------------------------ metrics
    store = {}
    n = 24000
    p, target, mask, prediction, border = synth_metric_inputs(n, 20)
    shape = (10, 40, 60)
    p3, t3, m3, d3, b3 = (a.reshape(shape) for a in (p, target, mask, prediction, border))
    store.update(p=p3, target=t3, mask=m3, prediction=d3, border=b3)
    to_eval = {'probabilities': p3.copy(), 'target': t3, 'prediction': d3, 'mask': m3, 'target_boarder': b3}
    to_eval = analysis.AddBackgroundProbabilities()(to_eval)
    store['prob2'] = to_eval['probabilities']
    to_eval = analysis.ToEntropy()(to_eval)
    store['uncertainty'] = to_eval['uncertainty']
    assert to_eval['uncertainty'].dtype == np.float64
    for with_mask in (False, True):
        for weighting in ('proportion', 'log_proportion', 'power_proportion', 'mean_proportion'):
            res = {}
            ev.EceBinaryNumpy(with_mask=with_mask, return_bins=True, bin_weighting=weighting)(to_eval, res)
            flatten_results('ece/mask%d/%s' % (with_mask, weighting), res, store)
    res = {}
    ev.EceBinaryNumpy(threshold_range=(0.2, 0.8))(to_eval, res)
    flatten_results('ece/range', res, store)
    res = {}
    ev.ComposeEvaluation([ev.DiceNumpy(), ev.ConfusionMatrix()])(to_eval, res)
    flatten_results('dice_cm', res, store)
    for th in R.SWEEP_THRESHOLDS:
        res = {}
        ev.UncertaintyAndCorrectionEvalNumpy(th)(to_eval, res)
        flatten_results('sweep/%s' % th, res, store)
        res = {}
        ev.UncertaintyErrorDiceNumpy(th, 'ue', with_mask=True)(to_eval, res)
        flatten_results('uedice_border/%s' % th, res, store)
        res = {}
        ev.UncertaintyErrorDiceNumpy(th)(to_eval, res)
        flatten_results('uedice/%s' % th, res, store)
    # degenerate inputs the reference handles in a defined way
    zeros = np.zeros(shape, dtype=np.uint8)
    deg = {'probabilities': np.stack([1 - p3, p3], -1), 'target': zeros, 'prediction': zeros, 'uncertainty': to_eval['uncertainty']}
    res = {}
    ev.UncertaintyAndCorrectionEvalNumpy(0.5)(deg, res)
    flatten_results('degenerate_empty/0.5', res, store)
    np.savez_compressed(os.path.join(OUT, 'metrics_golden.npz'), **store)
    print('wrote', sorted(os.listdir(OUT)))


if __name__ == '__main__':
    main()
