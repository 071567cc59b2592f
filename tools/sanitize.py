"""Tiny end-to-end pass over every kernel family for compute-sanitizer:

    compute-sanitizer --tool memcheck python tools/sanitize.py
    compute-sanitizer --tool initcheck python tools/sanitize.py

Small shapes (32 x 48 slices, T = 2, a 6-image chunk so that the chunk loop, partial tiles and the TMA out-of-bounds
paths all run): halo / wide / per-tap conv kernels, first conv, sigma head, features export, PostNet, aggregation,
fused histogram (single and batched), border mask, min / max, confidence preparation, Philox masks.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rcu_b200  # noqa: E402,F401
from rcu_b200 import metrics, model, steps, synth, tables  # noqa: E402

torch.set_grad_enabled(False)
sd = synth.random_unet_state_dict(in_channels=4, seed=20, sigma_out=True)
net = model.B200UNet(sd, in_channels=4, dropout=0.05, chunk_images=6)
post = model.B200PostNet(synth.random_postnet_state_dict())
x = torch.randn(3, 4, 32, 48)
out = net.forward_outputs(x, 3, dropout_mode=1, det_first=True, sigma=True, features=True, postnet=post)
again = net.forward_outputs(x, 3, dropout_mode=1, det_first=True, sigma=True, features=True, postnet=post)
for k in out:
    assert torch.equal(out[k], again[k]), k          # run-to-run determinism (also a cheap race detector for the launch overlap)
net.set_conv_impl(2)
out2 = net.forward_outputs(x, 3, dropout_mode=1, det_first=True)
net.set_conv_impl(0)
d = (out['logits'] - out2['logits']).abs().max().item()
assert d <= 2e-2 * out2['logits'].abs().max().item(), d   # wide kernel sums K in another order than the per-tap kernel
post(out['features'][0])
summ = steps.summarize(steps.LazyMultiProbabilities(out['logits'][1:]), do_mi=True, do_var=True, emit_prediction=True, emit_foreground=True)
fg, pred = summ['foreground'].reshape(-1), summ['prediction'].reshape(-1)
target = (torch.rand(fg.numel(), device=fg.device) < fg).to(torch.uint8)
mask = (torch.rand(fg.numel(), device=fg.device) < 0.5).to(torch.uint8)
metrics.eval_fused(fg, pred, target, mask)
metrics.eval_fused(fg, pred, target, None, n_subjects=3)
metrics.calibration_tables(fg, target, mask)
metrics.ue_tables(fg, pred, target, tables.SWEEP_THRESHOLDS)
metrics.confusion_counts(pred, target)
label = (torch.rand(5, 20, 24, device=fg.device) < 0.4).to(torch.uint8)
metrics.border_mask(label, 1, 1)
metrics.border_mask(label[0], 2, 3)
u = torch.rand(5, 20, 24, device=fg.device) * 2 + 0.25
metrics.confidence_to_foreground(u, label, rescale='subject')
metrics.philox_keep_scale(20, 0.05, net.site_channels, 0, 3, 0, 2)
torch.cuda.synchronize()
print('sanitize pass ok')
