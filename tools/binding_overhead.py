"""Host-side cost of one call through the two bindings of the C-ABI — ctypes (_lib.py) and the torch extension
(torch.ops.rcu_b200.*, csrc/torch_binding.cpp) — on the calls that are short enough for it to matter: the fused metric pass
(56 us of device time per subject) and the aggregation.  Tiny inputs make the loop host-bound, so wall time / calls is the
per-call launch path (argument marshalling, output allocation, the library's own host work, cudaLaunch).

    python tools/binding_overhead.py            (runs both routes in subprocesses: the routing is fixed at import time)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker():
    sys.path.insert(0, ROOT)
    import torch
    import rcu_b200  # noqa: F401
    from rcu_b200 import metrics, steps, tables, _torch_ext
    torch.set_grad_enabled(False)
    dev = torch.device('cuda:0')
    route = 'torch-extension' if _torch_ext.ops() is not None else 'ctypes'
    n = 4096
    p = torch.rand(n, device=dev)
    pred = (p > 0.5).to(torch.uint8)
    target = (torch.rand(n, device=dev) < p).to(torch.uint8)
    mask = torch.ones(n, dtype=torch.uint8, device=dev)
    bt = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
    logits = torch.randn((4, 1, 16, 16, 2), device=dev)

    def timeit(fn, reps=3000):
        for _ in range(200):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e6
    t_eval = timeit(lambda: metrics.eval_fused(p, pred, target, mask, 10, tables.SWEEP_THRESHOLDS, sync=False, break_table=bt))
    t_agg = timeit(lambda: steps.summarize(steps.LazyMultiProbabilities(logits), emit_prediction=True, emit_foreground=True))
    print('%-16s eval_fused %.1f us/call   summarize %.1f us/call   (host-bound loop, 4096 voxels / 256 pixels)' % (route, t_eval, t_agg))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'worker':
        worker()
    else:
        for binding in ('ctypes', 'torch'):
            env = dict(os.environ, RCU_B200_BINDING=binding)
            subprocess.run([sys.executable, os.path.abspath(__file__), 'worker'], env=env, check=False)
