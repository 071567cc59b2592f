"""The device-side assembler, subject step and metrics hook driven by the UNMODIFIED reference test loop
(common/trainloop/loops.py:204-235 `Test._test_batch`, hooks.py:116-151 `ReducedComposeTestLoopHook`), loaded through
oracle/ref_shim.py.  Runs only where the reference tree exists (the authoring container); CPU tensors stand in for CUDA
tensors — the assembler and the loop code are device agnostic, the kernels behind the hook are covered by the GPU tests."""
import pickle
import types

import numpy as np
import pytest
import torch

from oracle import ref_shim
from rcu_b200 import assembly, hooks as b200_hooks, nifti

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present on this box')


class IndexExpression:
    def __init__(self, expression):
        self.expression = expression


def test_unmodified_loop_with_device_assembler(tmp_path):
    ref_shim.load()
    import common.trainloop.loops as loops
    import common.trainloop.context as ctx
    import common.trainloop.hooks as ref_hooks
    import common.trainloop.steps as ref_steps

    g = torch.Generator().manual_seed(5)
    volumes = {0: torch.randn(9, 4, 8, 6, generator=g), 1: torch.randn(5, 4, 8, 6, generator=g)}
    weight = torch.randn(2, 4, generator=g)
    labels = {s: (torch.rand(v.shape[0], 8, 6, generator=g) < 0.3).numpy().astype(np.uint8) for s, v in volumes.items()}

    class PredictStep(ref_steps.BatchStep):              # a stand-in model call; the output protocol is what matters
        def __call__(self, batch_context, task_context, context):
            logits = torch.einsum('kc,nchw->nkhw', weight, batch_context.input['images'].float())
            batch_context.output['probabilities'] = torch.softmax(logits, 1)
            batch_context.output['dropped'] = logits                      # not in `entries`: must not be assembled

    seen = []

    class SubjectStep(ref_steps.SubjectStep):
        def __call__(self, subject_context, task_context, context):      # ExtractSubjectInfoStep's role: labels by direct extraction
            subject_context.subject_data['labels'] = labels[subject_context.subject_index]
            subject_context.metrics['n_slices'] = subject_context.subject_data['probabilities'].shape[0]

    class Recorder(ref_hooks.TestLoopHook):
        def on_test_subject_end(self, subject_context, task_context, context):
            seen.append((subject_context.subject_index, subject_context.subject_data, dict(subject_context.metrics)))

    evaluated = []

    class MetricsProbe(b200_hooks.DeviceMetricsHook):     # the real hook up to the kernel call (no GPU in this container)
        def evaluate(self, subject, p, prediction, target, mask=None):
            evaluated.append((subject, p, prediction, target))
            return {'subject': subject}

    metrics_hook = MetricsProbe()
    file_hook = nifti.AsyncNiftiWriteHook(out_dir=str(tmp_path))
    hook = ref_hooks.ReducedComposeTestLoopHook([Recorder(), metrics_hook, file_hook])
    test = loops.Test([PredictStep()], [SubjectStep()], assembly.DeviceSubjectAssembler(), entries=('probabilities',), convert_fn=None)
    samples = [(s, z) for s in sorted(volumes) for z in range(volumes[s].shape[0])]
    batch_size = 4
    n_batches = (len(samples) + batch_size - 1) // batch_size
    task_context = ctx.TaskContext(0, types.SimpleNamespace(nb_batches=n_batches), None)
    for b in range(n_batches):
        chunk = samples[b * batch_size:(b + 1) * batch_size]
        batch = {'images': torch.stack([volumes[s][z] for s, z in chunk]),
                 'subject_index': [s for s, _ in chunk],
                 'index_expr': [pickle.dumps(IndexExpression((z,))) for _, z in chunk],
                 'shape': [(volumes[s].shape[0], 8, 6) for s, _ in chunk]}
        test._test_batch(ctx.BatchContext(batch, b), task_context, None, hook)
    assert [s for s, _, _ in seen] == [0, 1]
    for subject, data, subject_metrics in seen:
        assert set(data) == {'probabilities', 'labels'} and torch.is_tensor(data['probabilities'])
        expected = torch.softmax(torch.einsum('kc,nchw->nkhw', weight, volumes[subject]), 1).permute(0, 2, 3, 1)
        assert torch.allclose(data['probabilities'], expected, rtol=0, atol=1e-6)   # einsum batches differ in the last ulp
        assert subject_metrics == {'n_slices': volumes[subject].shape[0]}
    assert task_context.history.get_entries('n_slices', 'subject_metrics') == [9, 5]
    # the metrics hook saw, per subject, what WriteHook would have stored: foreground p and the argmax prediction
    assert [e[0] for e in evaluated] == [0, 1] and [r['subject'] for r in metrics_hook.rows] == [0, 1]
    for subject, p, prediction, target in evaluated:
        probs = torch.softmax(torch.einsum('kc,nchw->nkhw', weight, volumes[subject]), 1)
        assert torch.allclose(p, probs[:, 1], rtol=0, atol=1e-6) and p.is_contiguous()
        assert (prediction != (probs[:, 1] > probs[:, 0]).to(torch.uint8)).sum().item() <= 1
        assert np.array_equal(target, labels[subject])
    # ... and the file hook wrote what WriteHook writes (bin-dl/brats_test_default.py:96-108), in the background
    hook.on_test_end(task_context, None)
    for subject in (0, 1):
        p_file, _ = nifti.read_nifti(str(tmp_path / '{}_probabilities.nii.gz'.format(subject)))
        d_file, _ = nifti.read_nifti(str(tmp_path / '{}_prediction.nii.gz'.format(subject)))
        assert p_file.dtype == np.float32 and d_file.dtype == np.uint8
        assert np.array_equal(p_file, [e for e in evaluated if e[0] == subject][0][1].numpy())
        assert np.array_equal(d_file, [e for e in evaluated if e[0] == subject][0][2].numpy())


def test_2d_assembler_in_the_unmodified_loop():
    ref_shim.load()
    import common.trainloop.loops as loops
    import common.trainloop.context as ctx
    import common.trainloop.hooks as ref_hooks
    import common.trainloop.steps as ref_steps

    class PredictStep(ref_steps.BatchStep):
        def __call__(self, batch_context, task_context, context):
            batch_context.output['probabilities'] = torch.softmax(batch_context.input['images'].float()[:, :2], 1)

    seen = {}

    class Recorder(ref_hooks.TestLoopHook):
        def on_test_subject_end(self, subject_context, task_context, context):
            seen[subject_context.subject_index] = subject_context.subject_data['probabilities']

    test = loops.Test([PredictStep()], [], assembly.DeviceSubject2dAssembler(), entries=None, convert_fn=None)
    images = torch.randn(5, 3, 8, 8, generator=torch.Generator().manual_seed(2))
    task_context = ctx.TaskContext(0, types.SimpleNamespace(nb_batches=2), None)
    for b, sl in enumerate((slice(0, 3), slice(3, 5))):
        batch = {'images': images[sl], 'subject_index': list(range(sl.start, sl.stop))}
        test._test_batch(ctx.BatchContext(batch, b), task_context, None, Recorder())
    assert sorted(seen) == [0, 1, 2, 3, 4]
    for i in range(5):
        assert torch.allclose(seen[i], torch.softmax(images[i:i + 1, :2], 1)[0].permute(1, 2, 0), rtol=0, atol=1e-6)


def test_all_entries_with_the_extra_maps_on_non_square_slices():
    """`entries=None` (bin-dl/brats_test_ensemble.py:63, isic_test_default.py:54-55): the unmodified loop applies
    th.channel_to_end to EVERY output it assembles (loops.py:210-216), so the extra 'prediction' / 'foreground' maps of
    the fused summary must carry a channel dimension like all reference outputs — (N, 1, H, W), the shape
    steps.MultiPredictionSummary emits.  Non-square slices: a missing channel dimension would transpose or fail."""
    ref_shim.load()
    import common.trainloop.loops as loops
    import common.trainloop.context as ctx
    import common.trainloop.hooks as ref_hooks
    import common.trainloop.steps as ref_steps

    g = torch.Generator().manual_seed(9)
    volume = torch.randn(6, 4, 8, 6, generator=g)
    weight = torch.randn(2, 4, generator=g)
    labels = (torch.rand(6, 8, 6, generator=g) < 0.3).numpy().astype(np.uint8)

    class SummaryLikeStep(ref_steps.BatchStep):          # the output protocol of McPredictStep + MultiPredictionSummary(emit_*=True)
        def __call__(self, batch_context, task_context, context):
            probs = torch.softmax(torch.einsum('kc,nchw->nkhw', weight, batch_context.input['images'].float()), 1)
            batch_context.output['probabilities'] = probs
            batch_context.output['entropy'] = -(probs * probs.log()).sum(1, keepdim=True)
            batch_context.output['prediction'] = (probs[:, 1] > probs[:, 0]).to(torch.uint8).unsqueeze(1)
            batch_context.output['foreground'] = probs[:, 1].unsqueeze(1).contiguous()

    class SubjectStep(ref_steps.SubjectStep):
        def __call__(self, subject_context, task_context, context):
            subject_context.subject_data['labels'] = labels

    evaluated = []

    class MetricsProbe(b200_hooks.DeviceMetricsHook):
        def evaluate(self, subject, p, prediction, target, mask=None):
            evaluated.append((p, prediction))
            return {'subject': subject}

    seen = {}

    class Recorder(ref_hooks.TestLoopHook):
        def on_test_subject_end(self, subject_context, task_context, context):
            seen.update(subject_context.subject_data)

    hook = ref_hooks.ReducedComposeTestLoopHook([Recorder(), MetricsProbe(probability_entry='foreground')])
    test = loops.Test([SummaryLikeStep()], [SubjectStep()], assembly.DeviceSubjectAssembler(), entries=None, convert_fn=None)
    task_context = ctx.TaskContext(0, types.SimpleNamespace(nb_batches=2), None)
    for b, sl in enumerate((slice(0, 4), slice(4, 6))):
        n = sl.stop - sl.start
        batch = {'images': volume[sl], 'subject_index': [0] * n,
                 'index_expr': [pickle.dumps(IndexExpression((z,))) for z in range(sl.start, sl.stop)], 'shape': [(6, 8, 6)] * n}
        test._test_batch(ctx.BatchContext(batch, b), task_context, None, hook)
    probs = torch.softmax(torch.einsum('kc,nchw->nkhw', weight, volume), 1)
    assert tuple(seen['probabilities'].shape) == (6, 8, 6, 2) and tuple(seen['entropy'].shape) == (6, 8, 6, 1)
    assert tuple(seen['prediction'].shape) == (6, 8, 6, 1) and tuple(seen['foreground'].shape) == (6, 8, 6, 1)
    assert torch.allclose(seen['foreground'][..., 0], probs[:, 1], rtol=0, atol=1e-6)
    # the metrics hook squeezes the channel: what it evaluates is the (Z, H, W) foreground map and the emitted argmax
    (p, prediction), = evaluated
    assert tuple(p.shape) == (6, 8, 6) and tuple(prediction.shape) == (6, 8, 6)
    assert torch.allclose(p, probs[:, 1], rtol=0, atol=1e-6)
    assert (prediction != (probs[:, 1] > probs[:, 0]).to(torch.uint8)).sum().item() <= 1
