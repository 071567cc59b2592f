"""Per-op device time of a few chunks of the BraTS MC forward: python tools/op_times.py [n_slices]."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rcu_b200  # noqa: E402,F401
from rcu_b200 import model, synth  # noqa: E402

torch.set_grad_enabled(False)
n_slices = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sd = synth.random_unet_state_dict(in_channels=4, seed=20)
net = model.B200UNet(sd, in_channels=4, dropout=0.05, device='cuda:0', seed=20)
x = torch.randn((n_slices, 4, 240, 240), generator=torch.Generator().manual_seed(1)).cuda()
for _ in range(2):
    net.forward_samples(x, 21, dropout_mode=1, det_first=True)
torch.cuda.synchronize()
net.enable_timing(True)
net.read_timing()
reps = 3
for _ in range(reps):
    net.forward_samples(x, 21, dropout_mode=1, det_first=True)
ms, launches = net.read_timing()
ops = net.op_table()
chunks = launches[0] / reps
print('chunks/forward=%d (chunk_images=%d)' % (chunks, net.chunk_images))
tot = 0.0
for i, o in enumerate(ops):
    if launches[i] == 0:
        continue
    per = ms[i] / reps / chunks
    tot += per
    tf = 2.0 * o['macs_per_image'] * (n_slices * 21.0 / chunks) / (per * 1e-3) / 1e12 if o['macs_per_image'] else 0
    print('op %2d %-10s %3d->%3d %3dx%-3d  %.4f ms/chunk  %6.0f TF' % (i, o['kind'], o['c_in'], o['c_out'], o['h'], o['w'], per, tf))
print('total %.3f ms/chunk' % tot)
