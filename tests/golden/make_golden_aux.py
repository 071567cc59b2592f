"""Generates tests/golden/aux_golden.npz by running the UNMODIFIED reference (through oracle/ref_shim.py) on the
single-forward methods that reuse the conv engine: the aleatoric sigma head, the auxiliary
feature PostNet, the 5-channel auxiliary segmentation input, the border mask and the confidence -> probability
preparations.  Run from the repo root in the authoring container:

    python tests/golden/make_golden_aux.py

The step classes of these methods live in the reference's scripts (bin-dl/*.py); they are loaded from their files
with importlib, unmodified.
"""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, restate as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def load_script(name):
    path = os.path.join(ref_shim.REFERENCE_ROOT, 'bin-dl', name + '.py')
    spec = importlib.util.spec_from_file_location('refscript_' + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref_shim.load()
    import common.model.unet as unet
    import common.model.postnet as postnet
    import common.trainloop.context as ctx
    import common.utils.labelhelper as lh
    import rechun.eval.helper as helper
    torch.set_grad_enabled(False)
    warnings.simplefilter('ignore')
    store = {}
    g = torch.Generator().manual_seed(2)

    # ---- aleatoric (sigma_out) ----
    aleatoric = load_script('brats_test_aleatoric')
    cfg = R.UNetConfig(in_channels=4, sigma_out=True)
    torch.manual_seed(20)
    net = unet.UNet(cfg.nb_classes, cfg.in_channels, depth=cfg.depth, start_filters=cfg.start_filters, dropout=cfg.dropout,
                    sigma_out=True).eval()
    sd = net.state_dict()
    store['aleatoric/param_sum'] = np.float64(sum(v.double().sum().item() for v in sd.values()))
    store['aleatoric/sigma_head_weight'] = sd['conv_sigma.1.weight'].numpy().copy()
    sd = R.randomize_statistics(sd, 7)
    sd['conv_sigma.1.weight'] = sd['conv_sigma.1.weight'] * 6.0
    net.load_state_dict(sd)
    x = torch.randn(2, 4, 48, 64, generator=g)
    store['aleatoric/input'] = x.numpy()
    c = ctx.TorchTestContext('cpu')
    c.model = net
    for is_log in (False, True):
        bc = ctx.BatchContext({'images': x.clone()}, 0)
        aleatoric.AleatoricPredictStep(is_log_sigma=is_log)(bc, None, c)
        tag = 'aleatoric/log%d/' % is_log
        for k in ('logits', 'sigma', 'probabilities'):
            store[tag + k] = bc.output[k].numpy()

    # ---- auxiliary features + PostNet ----
    auxfeat = load_script('brats_test_auxiliary_feat')
    cfg = R.UNetConfig(in_channels=4)
    torch.manual_seed(20)
    seg = unet.UNet(cfg.nb_classes, cfg.in_channels, depth=cfg.depth, start_filters=cfg.start_filters, dropout=cfg.dropout,
                    provide_features=True).eval()
    seg.load_state_dict(R.randomize_statistics(seg.state_dict(), 7))
    torch.manual_seed(21)
    post = postnet.PostNet(32, 2).eval()
    psd = post.state_dict()
    store['auxfeat/postnet_param_sum'] = np.float64(sum(v.double().sum().item() for v in psd.values()))
    store['auxfeat/postnet_logits_weight'] = psd['conv_logits.weight'].numpy().copy()
    psd = R.randomize_statistics(psd, 9)
    psd['conv_logits.weight'] = psd['conv_logits.weight'] * 4.0
    post.load_state_dict(psd)
    x = torch.randn(2, 4, 48, 64, generator=g)
    store['auxfeat/input'] = x.numpy()
    c = ctx.TorchTestContext('cpu')
    c.model = post
    bc = ctx.BatchContext({'images': x.clone()}, 0)
    auxfeat.SegmentationPredictStep(seg)(bc, None, c)
    store['auxfeat/segm_probabilities'] = bc.output['segm_probabilities'].numpy()
    store['auxfeat/probabilities'] = bc.output['probabilities'].numpy()
    store['auxfeat/features'] = seg.features.numpy()

    # ---- auxiliary segmentation (5 input channels) ----
    auxsegm = load_script('brats_test_auxiliary_segm')
    cfg = R.UNetConfig(in_channels=5)
    torch.manual_seed(20)
    net = unet.UNet(cfg.nb_classes, cfg.in_channels, depth=cfg.depth, start_filters=cfg.start_filters, dropout=cfg.dropout).eval()
    net.load_state_dict(R.randomize_statistics(net.state_dict(), 7))
    x = torch.randn(2, 4, 48, 64, generator=g)
    labels = (torch.rand(2, 2, 48, 64, generator=g) < 0.3).long()
    store['auxsegm/input'] = x.numpy()
    store['auxsegm/labels'] = labels.numpy()
    c = ctx.TorchTestContext('cpu')
    c.model = net
    bc = ctx.BatchContext({'images': x.clone(), 'labels': labels.clone()}, 0)
    auxsegm.SegmentationPredictStep()(bc, None, c)
    for k in ('logits', 'probabilities', 'orig_prediction'):
        store['auxsegm/' + k] = bc.output[k].numpy()

    # ---- border mask (common/utils/labelhelper.py:12-20) ----
    np.bool = bool  # the reference predates numpy 1.24 (labelhelper.py:13)
    rng = np.random.default_rng(5)
    zz, yy, xx = np.mgrid[0:12, 0:40, 0:36]
    blob = ((zz - 5.5) ** 2 / 16 + (yy - 18) ** 2 / 90 + (xx - 20) ** 2 / 60) < 1
    blob |= ((zz - 9) ** 2 + (yy - 30) ** 2 + (xx - 8) ** 2) < 10
    blob ^= rng.random(blob.shape) < 0.02          # speckle: isolated voxels and holes
    blob[0, :3, :3] = True                         # touches the volume corner
    store['border/label'] = blob.astype(np.uint8)
    for d_in, d_out in ((1, 1), (2, 1), (1, 2), (3, 3), (0, 1)):
        dist, mask = lh.boarder_mask(blob.astype(np.uint8), d_in, d_out)
        store['border/mask_%d_%d' % (d_in, d_out)] = mask
    store['border/dist'] = dist
    dist2, mask2 = lh.boarder_mask(blob[3].astype(np.uint8), 1, 1)   # a 2-D map (ISIC)
    store['border/mask2d_1_1'] = mask2

    # ---- confidence / sigma -> pseudo probabilities (rechun/eval/helper.py:7-22) ----
    u = rng.random((6, 20, 24)).astype(np.float32) * 3.0 + 0.5
    pred = (rng.random(u.shape) < 0.4).astype(np.uint8)
    resc = helper.rescale_uncertainties(u, float(u.min()), float(u.max()))
    store['prep/uncertainty'] = u
    store['prep/prediction'] = pred
    store['prep/rescaled'] = resc
    store['prep/foreground'] = helper.uncertainty_to_foreground_probabilities(resc.copy(), pred)
    np.savez_compressed(os.path.join(OUT, 'aux_golden.npz'), **store)
    print('wrote aux_golden.npz with', len(store), 'entries,', os.path.getsize(os.path.join(OUT, 'aux_golden.npz')), 'bytes')


if __name__ == '__main__':
    main()
