"""Phase timeline of the shared-atomic histogram kernel (debug build with -DRCU_ATOM_TRACE=1): per block the %globaltimer at
entry, after the prologue (tables built), after the streaming loop, after the block fold, after the fence, after the ticket,
at the last block's exit.  python tools/hist_trace.py  (builds nothing: run `python -c "import rcu_b200.build as b;
b.build_variant('trace', ['RCU_ATOM_TRACE=1'])"` first, on the build host)."""
import ctypes
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib_path = os.path.join(ROOT, 'reliability-challenges-uncertainty_b200', 'build', 'variants', 'librcu_b200_%s.so' % (sys.argv[1] if len(sys.argv) > 1 else 'trace'))
os.environ['RCU_B200_LIB'] = lib_path
os.environ['RCU_B200_BINDING'] = 'ctypes'
sys.path.insert(0, ROOT)
import torch
import rcu_b200  # noqa
from rcu_b200 import metrics, tables, _lib

torch.set_grad_enabled(False)
dev = torch.device('cuda:0')
bt = tables.uncertainty_break_table(tables.SWEEP_THRESHOLDS)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
L = _lib.lib()
fn = L.rcu_debug_atom_trace if hasattr(L, 'rcu_debug_atom_trace') else ctypes.CDLL(lib_path).rcu_debug_atom_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
for vps in (148 * 4096, 155 * 240 * 240):
    n = vps
    g = torch.Generator(device=dev).manual_seed(3)
    p = torch.rand(n, device=dev, generator=g)
    target = (torch.rand(n, device=dev, generator=g) < p).to(torch.uint8)
    pred = (p > 0.5).to(torch.uint8)
    mask = (torch.rand(n, device=dev, generator=g) < 0.25).to(torch.uint8)
    for rep in range(4):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if os.environ.get('TRACE_WARM'):   # the traced call runs right behind an identical one: kernel code and tables warm in L2
            metrics.eval_fused(p, pred, target, mask, 10, tables.SWEEP_THRESHOLDS, n_subjects=1, sync=False, break_table=bt)
        a.record()
        metrics.eval_fused(p, pred, target, mask, 10, tables.SWEEP_THRESHOLDS, n_subjects=1, sync=False, break_table=bt)
        b.record()
        b.synchronize()
        tr = np.zeros(2048 * 8, dtype=np.uint64)
        fn(tr.ctypes.data, tr.size)
        clk = tr[2040 * 8:2040 * 8 + 16].astype(np.int64)
        tr = tr.reshape(2048, 8).astype(np.int64)[:2040]
        used = tr[:, 0] > 0
        t = tr[used]
        t0 = t[:, 0].min()
        rel = (t[:, :7] - t0) / 1e3
        last = np.argmax(t[:, 6])
        print('vps %8d rep %d: event %.1f us, blocks %d | entry spread %.1f | prologue end (median/max) %.1f/%.1f | loop end %.1f/%.1f | fold end %.1f/%.1f | '
              'fence %.1f/%.1f | ticket %.1f/%.1f | exit of last %.1f'
              % (vps, rep, a.elapsed_time(b) * 1e3, used.sum(), rel[:, 0].max(), np.median(rel[:, 1]), rel[:, 1].max(), np.median(rel[:, 2]), rel[:, 2].max(),
                 np.median(rel[:, 3]), rel[:, 3].max(), np.median(rel[:, 4]), rel[:, 4].max(), np.median(rel[:, 5]), rel[:, 5].max(), rel[last, 6]), flush=True)
        print('   block 0 thread 0 clocks since entry:', [int(c - clk[0]) for c in clk[:15]], flush=True)
