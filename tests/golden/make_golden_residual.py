"""Fixture for residual=True nets (ConvResidualBlock, common/model/unet.py:42-60): the UNMODIFIED reference UNet built with
residual=True under torch.manual_seed(20), statistics randomised like the other goldens, run on one small input —
deterministic logits, features, and one MC-dropout sample with the Philox keep masks injected through forward hooks.
Pins oracle/restate.py's residual branch (init order, forward) and is the reference side of the GPU parity test.

    python tests/golden/make_golden_residual.py          (authoring container only: reads /root/reference)
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, restate as R  # noqa: E402


def main():
    ref_shim.load()
    import common.model.unet as unet
    torch.set_grad_enabled(False)
    cfg = R.UNetConfig(in_channels=4, residual=True)
    torch.manual_seed(20)
    net = unet.UNet(cfg.nb_classes, cfg.in_channels, depth=cfg.depth, start_filters=cfg.start_filters, dropout=cfg.dropout,
                    residual=True, provide_features=True).eval()
    sd = net.state_dict()
    store = {'param_sum': np.float64(sum(v.double().sum().item() for v in sd.values())),
             'param_abs_sum': np.float64(sum(v.double().abs().sum().item() for v in sd.values())),
             'keys': np.array(list(sd.keys())),
             'residual0_weight': sd['down_convs.0.block.residual.weight'].numpy().copy(),
             'residual_last_bias': sd['up_convs.3.block.residual.bias'].numpy().copy()}
    net.load_state_dict(R.randomize_statistics(sd, 7))
    x = torch.randn(2, 4, 48, 64, generator=torch.Generator().manual_seed(1))
    store['input'] = x.numpy()
    logits = net(x)
    store['logits'] = logits.numpy()
    store['features'] = net.features.numpy()
    # one stochastic pass with injected masks (th.set_dropout_mode semantics: only the Dropout2d modules in train mode)
    masks = R.philox_keep_masks(cfg, 20, 0, 0, x.shape[0])
    drops = [m for m in net.modules() if isinstance(m, nn.Dropout2d)]
    assert len(drops) == len(masks) == 19

    def make_hook(site):
        def hook(mod, inp, out):
            return inp[0] * (masks[site].float() / (1 - mod.p))[:, :, None, None]
        return hook
    handles = [d.register_forward_hook(make_hook(i)) for i, d in enumerate(drops)]
    for d in drops:
        d.train()
    store['mc_logits'] = net(x).numpy()
    for h in handles:
        h.remove()
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'residual_golden.npz'), **store)
    print('wrote residual_golden.npz', {k: getattr(v, 'shape', None) for k, v in store.items()})


if __name__ == '__main__':
    main()
