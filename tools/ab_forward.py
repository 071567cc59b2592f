"""A/B of library builds on ONE box: device time of the BraTS MC forward (155 slices, T = 20 + the weight-scaling pass) per
variant, interleaved so that clock / power drift hits every variant alike.

    python tools/ab_forward.py [rounds=2]       variants: the default library and build/variants/librcu_b200_*.so
"""
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import rcu_b200  # noqa: F401
    from rcu_b200 import model, synth
    torch.set_grad_enabled(False)
    net = model.B200UNet(synth.random_unet_state_dict(in_channels=4, seed=20), in_channels=4, dropout=0.05, device='cuda:0', seed=20)
    if os.environ.get('RCU_AB_DEDUP') == '0':
        net.set_first_layer_dedup(False)
    x = torch.randn((155, 4, 240, 240), generator=torch.Generator().manual_seed(1)).cuda()
    for i in range(3):
        net.forward_samples(x, 21, dropout_mode=1, det_first=True, slice_index0=i * 155)
    torch.cuda.synchronize()
    ts = []
    for i in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        net.forward_samples(x, 21, dropout_mode=1, det_first=True, slice_index0=(10 + i) * 155)
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    net.enable_timing(True)
    net.read_timing()
    net.forward_samples(x, 21, dropout_mode=1, det_first=True, slice_index0=0)
    ms, _ = net.read_timing()
    sel = {i: round(float(ms[i]), 2) for i in (0, 1, 4, 7, 13, 15, 18, 20, 21, 23, 24, 25, 26)}
    print('%-28s forward %.2f ms (min %.2f)  ops %s' % (os.environ.get('RCU_AB_NAME', 'default'), float(np.median(ts)), min(ts), sel), flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'worker':
        worker()
    else:
        rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
        libs = [('default', None), ('default_nodedup', None), ('default_tt2', None)] + [(os.path.basename(p)[len('librcu_b200_'):-3], p) for p in
                                      sorted(glob.glob(os.path.join(ROOT, 'reliability-challenges-uncertainty_b200', 'build', 'variants', '*.so')))]
        for _ in range(rounds):
            for name, path in libs:
                env = dict(os.environ, RCU_AB_NAME=name, RCU_B200_BINDING='ctypes')
                if name.endswith('_nodedup'):
                    env['RCU_AB_DEDUP'] = '0'
                if name.endswith('_tt2'):
                    env['RCU_HALO_TT'] = '2'     # two tiles per issue turn where the shared-memory ring is deep enough
                if path:
                    env['RCU_B200_LIB'] = path
                subprocess.run([sys.executable, os.path.abspath(__file__), 'worker'], env=env, check=False)
