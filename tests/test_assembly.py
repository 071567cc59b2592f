"""Device-side subject assembly: the assembler protocol of common/trainloop/loops.py:204-235 on
tensors.  The CPU tests drive it exactly the way `Test._test_batch` does (channel_to_end view of an NCHW output,
batches of 32 that straddle subjects, pickled index expressions); the GPU test runs prediction -> assembly ->
subject evaluation without a host copy and compares with the host route."""
import pickle

import numpy as np
import pytest
import torch

from rcu_b200 import assembly


class IndexExpression:  # stand-in for pymia.data.indexexpression.IndexExpression (only `.expression` is read)
    def __init__(self, expression):
        self.expression = expression


def _channel_to_end(t):  # common/utils/torchhelper.py:10-11
    return t.permute(0, *range(2, t.dim()), 1)


def _run_loop(assembler, volumes, batch_size, pickled, device='cpu', order=None):
    """volumes: {subject_index: (Z, C, H, W) tensor}.  Returns [(subject, assembled dict)] in completion order."""
    samples = [(s, z) for s in (order or sorted(volumes)) for z in range(volumes[s].shape[0])]
    done = []
    n_batches = (len(samples) + batch_size - 1) // batch_size
    for b in range(n_batches):
        chunk = samples[b * batch_size:(b + 1) * batch_size]
        out = torch.stack([volumes[s][z] for s, z in chunk]).to(device)
        expr = [IndexExpression((z,)) for _, z in chunk]
        batch = {'subject_index': [s for s, _ in chunk],
                 'index_expr': [pickle.dumps(e) for e in expr] if pickled else expr,
                 'shape': [tuple(volumes[s].shape[0:1]) + tuple(volumes[s].shape[2:]) for s, _ in chunk]}
        to_assemble = {'probabilities': _channel_to_end(out), 'entropy': _channel_to_end(out[:, :1])}
        assembler.add_batch(to_assemble, batch, last_batch=b == n_batches - 1)
        for subject in list(assembler.subjects_ready):
            done.append((subject, assembler.get_assembled_subject(subject)))
    return done


@pytest.mark.parametrize('pickled', [False, True])
@pytest.mark.parametrize('batch_size', [1, 4, 7, 32])
def test_slices_are_assembled_into_subject_volumes(batch_size, pickled):
    g = torch.Generator().manual_seed(3)
    volumes = {0: torch.rand(9, 2, 6, 5, generator=g), 1: torch.rand(13, 2, 6, 5, generator=g), 5: torch.rand(4, 2, 6, 5, generator=g)}
    done = _run_loop(assembly.DeviceSubjectAssembler(), volumes, batch_size, pickled)
    assert [s for s, _ in done] == [0, 1, 5]
    for subject, data in done:
        assert tuple(data['probabilities'].shape) == (volumes[subject].shape[0], 6, 5, 2)
        assert tuple(data['entropy'].shape) == (volumes[subject].shape[0], 6, 5, 1)
        assert torch.equal(data['probabilities'], volumes[subject].permute(0, 2, 3, 1))
        assert torch.equal(data['entropy'], volumes[subject][:, :1].permute(0, 2, 3, 1))


def test_subject_becomes_ready_when_the_next_one_starts_or_on_the_last_batch():
    a = assembly.DeviceSubjectAssembler()
    v = torch.arange(2 * 3 * 4, dtype=torch.float32).reshape(2, 3, 4)

    def batch(subjects, zs):
        return {'subject_index': subjects, 'index_expr': [IndexExpression((z,)) for z in zs], 'shape': [(2, 3, 4)] * len(zs)}
    a.add_batch({'x': v[:1]}, batch([0], [0]))
    assert a.subjects_ready == set()
    a.add_batch({'x': torch.stack([v[1], v[0]])}, batch([0, 1], [1, 0]))
    assert a.subjects_ready == {0}
    got = a.get_assembled_subject(0)
    assert torch.equal(got['x'], v) and a.subjects_ready == set()
    a.add_batch({'x': v[1:]}, batch([1], [1]), last_batch=True)
    assert a.subjects_ready == {1}
    assert torch.equal(a.get_assembled_subject(1)['x'], v)
    with pytest.raises(ValueError):
        a.get_assembled_subject(7)


def test_general_index_expressions_and_bare_tensor_input():
    a = assembly.DeviceSubjectAssembler()
    vol = torch.rand(4, 6, 6)
    exprs = [(slice(0, 2), slice(0, 3)), (slice(2, 4), slice(0, 3)), (slice(0, 2), slice(3, 6)), (slice(2, 4), slice(3, 6))]
    patches = torch.stack([vol[e] for e in exprs])
    a.add_batch(patches, {'subject_index': [2] * 4, 'index_expr': [IndexExpression(e) for e in exprs], 'shape': [(4, 6, 6)] * 4},
                last_batch=True)
    assert torch.equal(a.get_assembled_subject(2), vol)   # '__prediction' convention: bare tensor in, bare tensor out


def test_missing_batch_entries_raise_like_the_reference_assembler():
    a = assembly.DeviceSubjectAssembler()
    with pytest.raises(ValueError):
        a.add_batch({'x': torch.zeros(1, 2, 2)}, {'subject_index': [0], 'shape': [(1, 2, 2)]})
    with pytest.raises(ValueError):
        a.add_batch({'x': np.zeros((1, 2, 2))}, {'subject_index': [0], 'index_expr': [IndexExpression((0,))], 'shape': [(1, 2, 2)]})
    with pytest.raises(ValueError):  # slice does not fit the subject
        a.add_batch({'x': torch.zeros(1, 3, 2)}, {'subject_index': [0], 'index_expr': [IndexExpression((0,))], 'shape': [(1, 2, 2)]})


def test_2d_assembler_every_sample_is_a_subject():
    a = assembly.DeviceSubject2dAssembler()
    out = torch.rand(3, 2, 4, 4)
    a.add_batch({'probabilities': _channel_to_end(out)}, {'subject_index': [4, 5, 6]})
    assert a.subjects_ready == {4, 5, 6}
    for i, s in enumerate((4, 5, 6)):
        assert torch.equal(a.get_assembled_subject(s)['probabilities'], out[i].permute(1, 2, 0))
    assert a.subjects_ready == set()


@pytest.mark.gpu
def test_device_loop_equals_host_route():
    """steps -> DeviceSubjectAssembler -> EvalSubjectStep + DeviceMetricsHook, all on CUDA tensors, against the same
    subject evaluated from host numpy arrays (the reference's route)."""
    from rcu_b200 import hooks, steps

    class Ctx:
        def __init__(self, index, data):
            self.subject_index, self.subject_data, self.metrics = index, data, {}

    rng = np.random.default_rng(11)
    z, h, w = 37, 48, 32
    logits = torch.from_numpy(rng.normal(0, 2.5, size=(z, 2, h, w)).astype(np.float32)).cuda()
    prob = torch.softmax(logits, 1)
    labels = (rng.random((z, h, w)) < prob[:, 1].cpu().numpy()).astype(np.uint8)
    brain = rng.random((z, h, w)) < 0.4
    done = _run_loop(assembly.DeviceSubjectAssembler(), {3: prob}, 32, True, device='cuda')
    (subject, data), = done
    assert data['probabilities'].is_cuda and torch.equal(data['probabilities'], prob.permute(0, 2, 3, 1))
    rows = []
    for route in ('device', 'host'):
        d = {'probabilities': data['probabilities'] if route == 'device' else data['probabilities'].cpu().numpy(),
             'labels': torch.from_numpy(labels).cuda() if route == 'device' else labels,
             'mask': torch.from_numpy(brain).cuda() if route == 'device' else brain}
        c = Ctx(subject, d)
        steps.EvalSubjectStep()(c, None, None)
        hook = hooks.DeviceMetricsHook(mask_entry='mask')
        hook.on_test_subject_end(c, None, None)
        rows.append((c.metrics, hook.rows[0]))
    (m_dev, r_dev), (m_host, r_host) = rows
    assert m_dev == m_host and 0.0 < m_dev['dice'] < 1.0
    assert r_dev['ece'] == r_host['ece'] and r_dev['dice'] == r_host['dice'] == m_dev['dice']
    for th in r_dev['sweep']:
        assert r_dev['sweep'][th] == r_host['sweep'][th]
    pred = np.argmax(prob.permute(0, 2, 3, 1).cpu().numpy(), -1)
    tp = int(((pred == 1) & (labels == 1)).sum())
    assert r_dev['tp'] == tp
