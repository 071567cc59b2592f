"""A/B of the fused ECE + U-E histogram kernels on ONE box: the shared-atomic kernel (default) against the private-column
kernels (RCU_HIST_ATOM=0) and table sizes, on U-shaped iid maps (Beta(0.3, 0.3) p, Bernoulli(p) target, 25 % mask: the
worst case for data-dependent table reads) and on uniform p.  Every variant runs in its own process (the switches are read
once); tables must be identical across variants, float64 confidence sums equal to 1e-12.

    python tools/hist_ab.py [out.json]
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = [
    ('atom', {}),
    ('lut', {'RCU_HIST_ATOM': '0'}),
    ('atom_bits8', {'RCU_HIST_ATOM_BITS': '8'}),
    ('atom_bits9', {'RCU_HIST_ATOM_BITS': '9'}),
]


def worker():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import rcu_b200  # noqa: F401
    from rcu_b200 import metrics, tables
    torch.set_grad_enabled(False)
    dev = torch.device('cuda:0')
    bt = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    out = {}

    def timeit(fn, reps=12):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts)), float(min(ts))

    torch.manual_seed(5)
    torch.cuda.manual_seed(5)
    for shape in ('beta', 'uniform'):
        for S, vps in ((1, 148 * 4096), (1, 155 * 240 * 240), (50, 155 * 240 * 240)):
            n = S * vps
            g = torch.Generator(device=dev).manual_seed(3)
            if shape == 'beta':
                conc = torch.full((n,), 0.3, device=dev)
                ga = torch._standard_gamma(conc)
                gb = torch._standard_gamma(conc)
                p = (ga / (ga + gb)).clamp_(0, 1).nan_to_num_(0.5).float()
                del ga, gb, conc
            else:
                p = torch.rand(n, device=dev, generator=g)
            target = (torch.rand(n, device=dev, generator=g) < p).to(torch.uint8)
            pred = (p > 0.5).to(torch.uint8)
            mask = (torch.rand(n, device=dev, generator=g) < 0.25).to(torch.uint8)
            for with_mask in (True, False):
                m = mask if with_mask else None
                fn = lambda: metrics.eval_fused(p, pred, target, m, 10, tables.SWEEP_THRESHOLDS, n_subjects=S, sync=False, break_table=bt)
                med, mn = timeit(fn)
                r = metrics.eval_fused(p, pred, target, m, 10, tables.SWEEP_THRESHOLDS, n_subjects=S, break_table=bt)
                key = '%s_S%d_%s%s' % (shape, S, 'mask' if with_mask else 'nomask', '_tiny' if vps < 10 ** 6 else '')
                nb = 10
                out[key] = {'ms': med, 'ms_min': mn, 'us_per_subject': med * 1e3 / S, 'gbps': 7.0 * n / med / 1e6 if with_mask else 6.0 * n / med / 1e6,
                            'ints': [int(np.asarray(r[0]).sum()), int(np.asarray(r[1]).sum()), int(np.asarray(r[3]).sum()), int(np.asarray(r[4]).sum()),
                                     int((np.asarray(r[0])[:, :nb] * np.arange(1, nb + 1)).sum()), int((np.asarray(r[3]).reshape(S, -1) * np.arange(1, np.asarray(r[3]).reshape(S, -1).shape[1] + 1)).sum())],
                            'conf': [float(x) for x in np.asarray(r[2])[:, :nb].sum(axis=0)]}
            del p, target, pred, mask
            torch.cuda.empty_cache()
    print('RESULT ' + json.dumps(out), flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'worker':
        worker()
        sys.exit(0)
    res = {}
    for name, env in VARIANTS:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), 'worker'], env=dict(os.environ, **env), capture_output=True, text=True)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith('RESULT ')]
        if r.returncode != 0 or not line:
            print(name, 'FAILED', r.stderr[-3000:])
            continue
        res[name] = json.loads(line[0][7:])
        for k, v in res[name].items():
            print('%-10s %-22s %8.4f ms (min %.4f)  %7.2f us/subject  %7.0f GB/s  frac %.3f' % (name, k, v['ms'], v['ms_min'], v['us_per_subject'], v['gbps'], v['gbps'] / 6449.1), flush=True)
    ref = res.get('lut')
    ok = True
    if ref:
        for name, r in res.items():
            for k, v in r.items():
                if v['ints'] != ref[k]['ints']:
                    ok = False
                    print('MISMATCH ints', name, k, v['ints'], ref[k]['ints'])
                for a, b in zip(v['conf'], ref[k]['conf']):
                    if abs(a - b) > 1e-12 * max(abs(b), 1e-300):
                        ok = False
                        print('MISMATCH conf', name, k, a, b)
    print('tables identical across variants:', ok)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], 'w'), indent=1)
