"""Shared helpers for the test-suites (synthetic inputs that mirror tests/golden/make_golden.py)."""
import numpy as np

GOLDEN_CONFIGS = {'brats': dict(in_channels=4), 'isic': dict(in_channels=3),
                  'center': dict(in_channels=4, dropout=0.5, dropout_center=4)}
SWEEP = (0.05, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95)


def synth_metric_inputs(n, seed, with_break_neighbours=None):
    rng = np.random.default_rng(seed)
    p = rng.beta(0.3, 0.3, size=n).astype(np.float32)
    base = np.array([0.0, 1e-45, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1 - 2 ** -24], dtype=np.float32)
    adv = np.concatenate([base, np.nextafter(base, np.float32(2))[:-2], np.nextafter(base, np.float32(-1))[1:]])
    if with_break_neighbours is not None:
        br = np.asarray(with_break_neighbours, dtype=np.float32)
        adv = np.concatenate([adv, br, np.nextafter(br, np.float32(0)), np.nextafter(br, np.float32(2))])
        adv = adv[(adv >= 0) & (adv <= 1)]
    adv = adv[:n]
    p[:len(adv)] = adv
    target = (rng.random(n) < p).astype(np.uint8)
    mask = rng.random(n) < 0.25
    prediction = (p > 0.5).astype(np.uint8)
    flip = rng.random(n) < 0.05
    prediction[flip] ^= 1
    border = rng.random(n) < 0.1
    return p, target, mask, prediction, border


def results_equal(got, exp, key, rtol=0.0):
    """Compare one result entry with its golden value (NaN == NaN, exact unless rtol given)."""
    g, e = np.asarray(got), np.asarray(exp)
    assert g.shape == e.shape, (key, g.shape, e.shape)
    if e.dtype.kind in 'biu':
        assert np.array_equal(g.astype(e.dtype), e), (key, g, e)
    else:
        assert np.allclose(g.astype(np.float64), e.astype(np.float64), rtol=rtol, atol=0.0, equal_nan=True), (key, g, e)
