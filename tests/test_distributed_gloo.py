"""World-size-2 gloo tests of the multi-GPU host logic (sharding arithmetic, exact table all-reduce, probability-sum
all-reduce followed by the summary maths) — the kernels themselves are covered by the -m gpu suites."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rcu_b200 import distributed as D
from rcu_b200 import tables
from oracle import restate as R
from helpers import synth_metric_inputs


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 20, 50, 155, 7750):
        for world in (1, 2, 3, 4, 8):
            sizes = D.shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
            edges = [D.shard_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
    assert D.shard_sizes(20, 8) == [3, 3, 3, 3, 2, 2, 2, 2]     # T=20 over 8 GPUs
    assert D.shard_sizes(10, 8) == [2, 2, 1, 1, 1, 1, 1, 1]     # 10 ensemble members over 8 GPUs
    with pytest.raises(ValueError):
        D.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # ---- (1) a subject whose slices span ranks: per-rank tables, exact integer all-reduce ----
        n = 40000
        p, target, mask, pred, _ = synth_metric_inputs(n, 11)
        lo, hi = D.shard_bounds(n, world, rank)
        cnt, pos, conf = R.calibration_tables(p[lo:hi], target[lo:hi], mask=mask[lo:hi])
        unc = R.normalized_entropy(R.add_background_probability(p[lo:hi]))
        j = (unc[:, None] > np.array(tables.SWEEP_THRESHOLDS)[None, :]).sum(1)
        t, d = target[lo:hi].astype(bool), pred[lo:hi].astype(bool)
        ue = np.array([[np.sum(r & (j == c)) for c in range(12)] for r in (t & d, ~t & ~d, ~t & d, t & ~d)], dtype=np.int64)
        tens = [torch.from_numpy(np.ascontiguousarray(a)) for a in (cnt.astype(np.int64), pos.astype(np.int64), conf, ue)]
        D.allreduce_metric_tables_(*tens)
        # ---- (2) MC samples split over ranks: sum of per-sample softmax, all-reduce, then mean/entropy ----
        T = 5
        g = torch.Generator().manual_seed(3)
        logits = torch.randn(T, 2, 2, 8, 8, generator=g) * 2
        probs = torch.softmax(logits, 2)
        t_lo, t_hi = D.shard_bounds(T, world, rank)
        sums = probs[t_lo:t_hi].sum(0) if t_hi > t_lo else torch.zeros_like(probs[0])
        D.allreduce_sum_(sums)
        rows = D.gather_rows([{'rank': rank, 'subject': s} for s in range(*D.shard_bounds(5, world, rank))])
        if rank == 0:
            torch.save({'tables': [x.numpy() for x in tens], 'sums': sums, 'rows': rows}, os.path.join(out_dir, 'r0.pt'))
    finally:
        dist.destroy_process_group()


def test_world_size_2_allreduce_paths(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(str(tmp_path), 'r0.pt'), weights_only=False)
    n = 40000
    p, target, mask, pred, _ = synth_metric_inputs(n, 11)
    cnt, pos, conf = R.calibration_tables(p, target, mask=mask)
    assert np.array_equal(got['tables'][0], cnt) and np.array_equal(got['tables'][1], pos.astype(np.int64))
    assert np.allclose(got['tables'][2], conf, rtol=1e-12, atol=0)
    unc = R.normalized_entropy(R.add_background_probability(p))
    for k, th in enumerate(tables.SWEEP_THRESHOLDS):
        exp = R.uncertainty_counts(pred.astype(bool), target.astype(bool), unc > th)
        assert tuple(int(v) for v in tables.counts_at_threshold(got['tables'][3], k)) == tuple(int(v) for v in exp)
    # ECE from the all-reduced tables == single-process ECE
    ece = tables.ece_from_tables(got['tables'][0], got['tables'][1], got['tables'][2], n_dim=1)
    assert np.isclose(ece, R.ece_binary(p, target, mask=mask)[0], rtol=1e-12)
    g = torch.Generator().manual_seed(3)
    logits = torch.randn(5, 2, 2, 8, 8, generator=g) * 2
    ref = R.summarize(torch.softmax(logits, 2))
    mean = got['sums'] / 5
    assert torch.allclose(mean, ref['probabilities'], rtol=0, atol=2e-7)
    assert torch.allclose(R.torch_entropy(mean, 1, True), ref['entropy'], rtol=0, atol=5e-7)
    assert [r['subject'] for r in got['rows']] == [0, 1, 2, 3, 4] and [r['rank'] for r in got['rows']] == [0, 0, 0, 1, 1]


def _cohort(n_subjects=6, n=6000):
    out = []
    for s in range(n_subjects):
        p, target, mask, pred, _ = synth_metric_inputs(n, 30 + s)
        out.append((np.clip(p * (0.6 + 0.08 * s), 0, 1).astype(np.float32), target, mask, pred))
    return out


def _subject_row(s, p, target, mask, pred):
    cnt, pos, conf = R.calibration_tables(p, target, mask=mask)
    unc = R.normalized_entropy(R.add_background_probability(p))
    return {'subject': s, 'count': cnt.astype(np.int64), 'positives': pos.astype(np.int64), 'conf': conf,
            'ece': R.ece_binary(p, target, mask=mask)[0], 'dice': R.dice(pred, target),
            'sweep': {th: R.uncertainty_and_correction(pred, target, unc, th) for th in tables.SWEEP_THRESHOLDS}}


def _report_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        cohort = _cohort()
        lo, hi = D.shard_bounds(len(cohort), world, rank)           # whole subjects per rank: no table all-reduce needed
        rows = D.gather_rows([_subject_row(s, *cohort[s]) for s in range(lo, hi)])
        if rank == 0:
            torch.save(rows, os.path.join(out_dir, 'rows.pt'))
    finally:
        dist.destroy_process_group()


def test_world_size_2_subject_sharded_report_tables(tmp_path):
    """Config-5 style sharding (whole subjects per rank): the gathered per-subject rows give the same
    data-set level ECE and best-threshold table as one process over all subjects."""
    world = 2
    mp.spawn(_report_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rows = torch.load(os.path.join(str(tmp_path), 'rows.pt'), weights_only=False)
    assert [r['subject'] for r in rows] == list(range(6))
    single = [_subject_row(s, *c) for s, c in enumerate(_cohort())]
    got = tables.dataset_vs_mean_subject_ece(*(np.stack([r[k] for r in rows]) for k in ('count', 'positives', 'conf')))
    exp = tables.dataset_vs_mean_subject_ece(*(np.stack([r[k] for r in single]) for k in ('count', 'positives', 'conf')))
    assert got == exp and np.isclose(got['ece'], np.mean([r['ece'] for r in single]), rtol=1e-12)
    # the pooled ECE equals the ECE of the concatenated voxels
    cohort = _cohort()
    pooled = R.ece_binary(np.concatenate([c[0] for c in cohort]), np.concatenate([c[1] for c in cohort]),
                          mask=np.concatenate([c[2] for c in cohort]))[0]
    assert np.isclose(got['ds_ece'], pooled, rtol=1e-12)
    a = tables.best_threshold_summary([r['sweep'] for r in rows], [r['ece'] for r in rows], [r['dice'] for r in rows])
    b = R.best_threshold_summary([r['sweep'] for r in single], [r['ece'] for r in single], [r['dice'] for r in single])
    assert a == b
