"""Property-based CPU tests (hypothesis) of the host logic around the kernels: the device assembler under arbitrary
batchings, the table reductions against the oracle on arbitrary count tables, the break table under arbitrary thresholds."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from rcu_b200 import assembly, tables
from oracle import restate as R


class IndexExpression:
    def __init__(self, expression):
        self.expression = expression


@settings(max_examples=40, deadline=None)
@given(sizes=st.lists(st.integers(1, 9), min_size=1, max_size=5), batch_size=st.integers(1, 11), seed=st.integers(0, 10 ** 6))
def test_assembler_reassembles_any_batching(sizes, batch_size, seed):
    g = torch.Generator().manual_seed(seed)
    volumes = [torch.rand(z, 2, 3, 4, generator=g) for z in sizes]
    samples = [(s, z) for s, v in enumerate(volumes) for z in range(v.shape[0])]
    a = assembly.DeviceSubjectAssembler()
    done = []
    n_batches = (len(samples) + batch_size - 1) // batch_size
    for b in range(n_batches):
        chunk = samples[b * batch_size:(b + 1) * batch_size]
        out = torch.stack([volumes[s][z] for s, z in chunk])
        batch = {'subject_index': [s for s, _ in chunk], 'index_expr': [IndexExpression((z,)) for _, z in chunk],
                 'shape': [(volumes[s].shape[0], 3, 4) for s, _ in chunk]}
        a.add_batch({'p': out.permute(0, 2, 3, 1)}, batch, last_batch=b == n_batches - 1)
        for s in sorted(a.subjects_ready):
            done.append((s, a.get_assembled_subject(s)['p']))
    assert [s for s, _ in done] == list(range(len(volumes)))
    for s, p in done:
        assert torch.equal(p, volumes[s].permute(0, 2, 3, 1))
    assert a.predictions == {} and a.subjects_ready == set()


@settings(max_examples=60, deadline=None)
@given(counts=st.lists(st.integers(0, 10 ** 6), min_size=10, max_size=10), seed=st.integers(0, 10 ** 6),
       weighting=st.sampled_from(['proportion', 'log_proportion', 'power_proportion', 'mean_proportion']))
def test_ece_from_tables_matches_oracle_on_any_table(counts, seed, weighting):
    rng = np.random.default_rng(seed)
    count = np.array(counts, dtype=np.int64)
    if count.sum() == 0:
        count[3] = 5
    if weighting == 'log_proportion':
        count[count == 1] = 2          # log(1) = 0 weights: the reference divides by their sum (0/0 when every bin holds 1)
    positives = (count * rng.random(10)).astype(np.int64)
    conf = count * (np.arange(10) + rng.random(10)) / 10.0
    bins = {}
    got = tables.ece_from_tables(count, positives, conf, weighting, n_dim=3, out_bins=bins)
    exp, exp_bins = R.ece_from_tables(count, positives.astype(np.float64), conf, weighting, 3)
    assert got == exp or (np.isnan(got) and np.isnan(exp))
    for k in exp_bins:
        assert np.array_equal(bins[k], exp_bins[k])
    full = tables.expand_bins(dict(bins, ece=got))
    assert len(full['bins_count']) == 10 and np.array_equal(full['bins_count'][count != 0], bins['bins_count'])
    assert full['bins_count'][count == 0].sum() == 0


@settings(max_examples=25, deadline=None)
@given(ths=st.lists(st.floats(0.01, 0.99), min_size=1, max_size=6, unique=True), seed=st.integers(0, 10 ** 6))
def test_break_table_classifies_like_the_reference_arithmetic(ths, seed):
    breaks, seg, order = tables.uncertainty_break_table(tuple(ths))
    rng = np.random.default_rng(seed)
    p = rng.random(4000).astype(np.float32)
    near = np.concatenate([np.nextafter(breaks, np.float32(0)), breaks, np.nextafter(breaks, np.float32(2))])
    p = np.concatenate([p, near[(near >= 0) & (near <= 1)]]).astype(np.float32)
    j = seg[np.searchsorted(breaks, p, side='right')]
    u = R.normalized_entropy(R.add_background_probability(p))
    assert np.array_equal(j, (u[:, None] > np.sort(np.array(ths))[None, :]).sum(1))
    assert sorted(order) == list(range(len(ths)))


@settings(max_examples=60, deadline=None)
@given(c=st.lists(st.integers(0, 5000), min_size=8, max_size=8))
def test_correction_results_follow_from_the_eight_counts(c):
    tpu, tnu, fpu, fnu = c[4:]
    tp, tn, fp, fn = c[0] + tpu, c[1] + tnu, c[2] + fpu, c[3] + fnu      # the *u counts are subsets
    r = tables.correction_results(tp, tn, fp, fn, tpu, tnu, fpu, fnu)
    # materialise maps with exactly these counts and run the reference arithmetic on them
    parts = [(1, 1, 0, tp - tpu), (0, 0, 0, tn - tnu), (1, 0, 0, fp - fpu), (0, 1, 0, fn - fnu),
             (1, 1, 1, tpu), (0, 0, 1, tnu), (1, 0, 1, fpu), (0, 1, 1, fnu)]
    pred = np.concatenate([np.full(n, d, dtype=np.uint8) for d, t, u, n in parts])
    target = np.concatenate([np.full(n, t, dtype=np.uint8) for d, t, u, n in parts])
    unc = np.concatenate([np.full(n, 0.9 if u else 0.1) for d, t, u, n in parts])
    if pred.size == 0:
        return
    exp = R.uncertainty_and_correction(pred, target, unc, 0.5)
    assert set(r) == set(exp)
    for k, v in exp.items():
        a, b = np.asarray(r[k]), np.asarray(v)
        assert (a == b) or (np.isnan(a.astype(float)) and np.isnan(b.astype(float))), (k, a, b)
