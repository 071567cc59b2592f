"""Device time of the adjacent-row kernels (one BraTS subject, 155 x 240 x 240): sigma-head forward, features export,
PostNet (fused on the bf16 workspace features / stand-alone on float32 NCHW), border mask, min/max, confidence
preparation — each with the bytes (or FLOPs) that bound it."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rcu_b200  # noqa: E402,F401
from rcu_b200 import metrics, model, synth  # noqa: E402

torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
Z, H, W = 155, 240, 240
vox = Z * H * W
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


x = torch.randn(Z, 4, H, W, device=dev)
sd = synth.random_unet_state_dict(in_channels=4, seed=20, sigma_out=True)
net = model.B200UNet(sd, in_channels=4, dropout=0.05)
base = timeit(lambda: net.forward_outputs(x, 1))
with_sigma = timeit(lambda: net.forward_outputs(x, 1, sigma=True))
print('deterministic forward, one subject: %.3f ms; with the sigma head: %.3f ms (+%.3f ms for conv_sigma.0 + 1x1)' % (base, with_sigma, with_sigma - base))
post = model.B200PostNet(synth.random_postnet_state_dict())
fused = timeit(lambda: net.forward_outputs(x, 1, postnet=post))
print('  + fused PostNet: %.3f ms (+%.3f ms; %.1f GFMA -> %.1f TFLOP/s fp32)' % (fused, fused - base, 3136 * vox / 1e9, 2 * 3136 * vox / (fused - base) / 1e9))
feat = timeit(lambda: net.forward_outputs(x, 1, features=True))
print('  + features export: %.3f ms (+%.3f ms; %.0f GB/s of 64 B read + 128 B written per pixel)' % (feat, feat - base, 192.0 * vox / (feat - base) / 1e6))
f32 = net.forward_outputs(x, 1, features=True)['features'][0]
alone = timeit(lambda: post(f32))
print('PostNet on float32 NCHW features: %.3f ms (%.0f GB/s of 136 B/pixel, %.1f TFLOP/s fp32)' % (alone, 136.0 * vox / alone / 1e6, 2 * 3136 * vox / alone / 1e9))
del f32

g = torch.Generator(device=dev).manual_seed(1)
zz, yy, xx = torch.meshgrid(torch.arange(Z, device=dev), torch.arange(H, device=dev), torch.arange(W, device=dev), indexing='ij')
label = ((((zz - 70) / 40.0) ** 2 + ((yy - 120) / 60.0) ** 2 + ((xx - 100) / 50.0) ** 2) < 1).to(torch.uint8)
for d in (1, 2, 3):
    ms = timeit(lambda: metrics.border_mask(label, d, d))
    print('border mask d=%d: %.4f ms (%.0f GB/s of 2 B/voxel)' % (d, ms, 2.0 * vox / ms / 1e6))
u = torch.rand(Z, H, W, device=dev, generator=g) * 3 + 0.5
ms = timeit(lambda: metrics.minmax(u))
print('min/max (incl. the 12-byte read-back): %.4f ms (%.0f GB/s of 4 B/voxel)' % (ms, 4.0 * vox / ms / 1e6))
ms = timeit(lambda: metrics.confidence_to_foreground(u, label, rescale=(0.5, 3.5)))
print('confidence -> foreground p (incl. the invalid-count read-back): %.4f ms (%.0f GB/s of 9 B/voxel)' % (ms, 9.0 * vox / ms / 1e6))
