"""Drop-in BatchSteps for the reference's test loop (common/trainloop/loops.py:204-223).

Same names, constructor arguments, call protocol `(batch_context, task_context, context)`, output keys, shapes,
dtypes and device placement as
  SegmentationPredictStep   common/trainloop/steps.py:69-90
  McPredictStep             rechun/dl/customsteps.py:10-39
  MultiPredictionSummary    rechun/dl/customsteps.py:42-71
  EnsemblePredictionStep    bin-dl/brats_test_ensemble.py:72-94 (ISIC twin bin-dl/isic_test_ensemble.py:73-95)
  AleatoricPredictStep      bin-dl/brats_test_aleatoric.py:51-73 (sigma_out nets)
  AuxiliaryFeatPredictStep  bin-dl/brats_test_auxiliary_feat.py:61-80 (`SegmentationPredictStep(test_model)` there)
  AuxiliarySegmPredictStep  bin-dl/brats_test_auxiliary_segm.py:48-69 (`SegmentationPredictStep()` there)
but the T+1 (or M) forwards run folded into one batch on the tcgen05 engine and softmax / mean / entropy are one
fused pass over the logits.  `batch_context.output['multi_probabilities']` is a LazyMultiProbabilities: the fused
summary consumes the logits behind it directly; anything else that touches it gets the real (T, N, C, H, W) tensor.
"""
import abc
import itertools
import os
import weakref

import torch

from . import _lib
from . import _torch_ext
from .model import B200UNet, B200PostNet

try:  # inside the reference tree the real protocol classes are used, so isinstance checks keep working
    import common.trainloop.steps as _ref_steps
    import common.trainloop.context as _ref_ctx
    _BatchStepBase = _ref_steps.BatchStep
    _CONTEXT_TYPES = (_ref_ctx.TorchTrainContext, _ref_ctx.TorchTestContext)
except Exception:  # standalone use: duck-typed contexts
    _ref_ctx = None
    _CONTEXT_TYPES = None

    class _BatchStepBase(abc.ABC):
        @abc.abstractmethod
        def __call__(self, batch_context, task_context, context) -> None:
            pass


def _type_error_msg(obj, expected_cls):
    # common/utils/messages.py:4-10
    exp = '({})'.format(','.join(e.__name__ for e in expected_cls)) if len(expected_cls) > 1 else expected_cls[0].__name__
    return 'expected type is "{}" but object is of type "{}"'.format(exp, obj.__class__.__name__)


def _check_context(context):
    if _CONTEXT_TYPES is not None:
        if not isinstance(context, _CONTEXT_TYPES):
            raise ValueError(_type_error_msg(context, _CONTEXT_TYPES))
    elif not (hasattr(context, 'model') and hasattr(context, 'device')):
        raise ValueError('expected type is "(TorchTrainContext,TorchTestContext)" but object is of type "{}"'.format(
            context.__class__.__name__))


_ENGINES = weakref.WeakKeyDictionary()   # reference module -> (weight fingerprint, converted engine); dies with the module


def _weights_fingerprint(module):
    """Changes whenever a parameter or buffer of `module` is rebound or written in place (load_state_dict / an optimizer
    step bump the tensors' version counters), so a cached conversion never outlives the weights it was folded from."""
    return tuple((t.data_ptr(), t._version) for t in itertools.chain(module.parameters(), module.buffers()))


def engine_for(model, device=None, seed=None):
    """The B200UNet behind a model: the model itself, or a cached conversion of a reference UNet (rebuilt when the
    module's weights change, e.g. context.load_from_checkpoint between two evaluations of one model object)."""
    if isinstance(model, B200UNet):
        return model
    fp = _weights_fingerprint(model)
    eng = _ENGINES.get(model)
    if eng is None or eng[0] != fp:
        kw = {} if seed is None else {'seed': seed}
        eng = (fp, B200UNet.from_reference(model, device=device, **kw))
        _ENGINES[model] = eng
    return eng[1]


def release_engines():
    """Drop every cached conversion (their device weights go with them; the shared activation workspace stays)."""
    _ENGINES.clear()
    _POSTNETS.clear()


# McPredictStep asks the engine for logit differences instead of logit pairs (RCU_LOGIT_DIFF=0: the pairs, for A/B and parity)
LOGIT_DIFF = os.environ.get('RCU_LOGIT_DIFF', '1') != '0'


def softmax_planar(logits_interleaved, diff=False):
    """(N, H, W, 2) interleaved logits — or (N, H, W) logit differences l0 - l1 with diff=True — -> (N, 2, H, W) probabilities
    through the aggregation kernel (1 sample)."""
    n, h, w = logits_interleaved.shape[:3]
    src = logits_interleaved.contiguous()
    out = torch.empty((n, 2, h, w), dtype=torch.float32, device=src.device)
    _lib.check(_lib.lib().rcu_aggregate(_lib.ptr(src), 3 if diff else 0, 1, n, h * w, _lib.ptr(out), None, None, None, None, None, None,
                                        _lib.current_stream()))
    return out


class _LazyTensor:
    """A stand-in that turns into the real tensor the moment anything tensor-like touches it."""

    _tensor = None

    def materialize(self):
        raise NotImplementedError

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        args = tuple(a.materialize() if isinstance(a, _LazyTensor) else a for a in args)
        return func(*args, **kwargs)

    def __getattr__(self, name):  # anything tensor-like falls through to the real tensor
        if name.startswith('__'):
            raise AttributeError(name)
        return getattr(self.materialize(), name)


class LazyWsProbabilities(_LazyTensor):
    """`ws_probabilities` of McPredictStep (rechun/dl/customsteps.py:23-25) before its softmax has run.  The fused
    MultiPredictionSummary computes it in its own launch and puts the real tensor into `batch_context.output`; any other
    consumer that touches it first gets it from a stand-alone softmax pass."""

    def __init__(self, logits, diff=False):
        self.logits = logits  # (N, H, W, 2) interleaved, float32 — or (N, H, W) differences l0 - l1 (diff=True)
        self.diff = diff

    @property
    def shape(self):
        n, h, w = self.logits.shape[:3]
        return torch.Size((n, 2, h, w))

    def materialize(self):
        if self._tensor is None:
            self._tensor = softmax_planar(self.logits, self.diff)
        return self._tensor


class LazyMultiProbabilities(_LazyTensor):
    """Stands in for the stacked per-sample probabilities (T, N, 2, H, W) without materialising them."""

    def __init__(self, logits, diff=False):
        self.logits = logits  # (T, N, H, W, 2) interleaved, float32 — or (T, N, H, W) differences l0 - l1 (diff=True)
        self.diff = diff
        self._tensor = None

    @property
    def shape(self):
        t, n, h, w = self.logits.shape[:4]
        return torch.Size((t, n, 2, h, w))

    def materialize(self):
        if self._tensor is None:
            t, n, h, w = self.logits.shape[:4]
            src = self.logits.contiguous()
            mean = torch.empty((n, 2, h, w), dtype=torch.float32, device=src.device)
            multi = torch.empty((t, n, 2, h, w), dtype=torch.float32, device=src.device)
            _lib.check(_lib.lib().rcu_aggregate(_lib.ptr(src), 3 if self.diff else 0, t, n, h * w, _lib.ptr(mean), None, None, None, None,
                                                None, _lib.ptr(multi), _lib.current_stream()))
            self._tensor = multi
        return self._tensor


class SegmentationPredictStep(_BatchStepBase):

    def __init__(self, has_labels=False, do_probs=False) -> None:
        super().__init__()
        self.has_labels = has_labels
        self.do_probs = do_probs

    def __call__(self, batch_context, task_context, context) -> None:
        _check_context(context)
        batch_context.input['images'] = batch_context.input['images'].float().to(context.device, non_blocking=True)
        if self.has_labels:
            batch_context.input['labels'] = batch_context.input['labels'].long().to(context.device)
        engine = engine_for(context.model, context.device)
        logits = engine.forward_samples(batch_context.input['images'], 1, dropout_mode=0)[0]
        batch_context.output['logits'] = logits.permute(0, 3, 1, 2)
        if self.do_probs:
            batch_context.output['probabilities'] = softmax_planar(logits)


class AleatoricPredictStep(_BatchStepBase):
    """One deterministic forward of a sigma_out net: 'logits', 'sigma' (= |out_sigma| or exp(out_sigma)) and
    'probabilities' = softmax(logits) (bin-dl/brats_test_aleatoric.py:51-73)."""

    def __init__(self, is_log_sigma=False) -> None:
        super().__init__()
        self.is_log_sigma = is_log_sigma

    def __call__(self, batch_context, task_context, context) -> None:
        _check_context(context)
        batch_context.input['images'] = batch_context.input['images'].float().to(context.device, non_blocking=True)
        engine = engine_for(context.model, context.device)
        out = engine.forward_outputs(batch_context.input['images'], 1, dropout_mode=0, sigma=True)
        logits, sigma = out['logits'][0], out['sigma'][0].permute(0, 3, 1, 2)
        batch_context.output['logits'] = logits.permute(0, 3, 1, 2)
        batch_context.output['sigma'] = sigma.exp_() if self.is_log_sigma else sigma.abs_()
        batch_context.output['probabilities'] = softmax_planar(logits)


_POSTNETS = weakref.WeakKeyDictionary()


def postnet_for(model, device=None):
    """The B200PostNet behind an auxiliary model: the model itself, or a cached conversion of a reference PostNet."""
    if isinstance(model, B200PostNet):
        return model
    fp = _weights_fingerprint(model)
    post = _POSTNETS.get(model)
    if post is None or post[0] != fp:
        post = (fp, B200PostNet.from_reference(model, device=device))
        _POSTNETS[model] = post
    return post[1]


class AuxiliaryFeatPredictStep(_BatchStepBase):
    """The auxiliary-feature method (bin-dl/brats_test_auxiliary_feat.py:61-80): the segmentation net's probabilities
    ('segm_probabilities') and the PostNet's probabilities on the segmentation net's features ('probabilities').  The
    PostNet runs inside the same engine call on the bf16 features in the workspace; the (N, 32, H, W) float32
    `features` tensor of the reference is never materialised."""

    def __init__(self, test_model) -> None:
        super().__init__()
        self.test_model = test_model

    def __call__(self, batch_context, task_context, context) -> None:
        _check_context(context)
        batch_context.input['images'] = batch_context.input['images'].float().to(context.device, non_blocking=True)
        engine = engine_for(self.test_model, context.device)
        post = postnet_for(context.model, context.device)
        out = engine.forward_outputs(batch_context.input['images'], 1, dropout_mode=0, postnet=post)
        batch_context.output['segm_probabilities'] = softmax_planar(out['logits'][0])
        batch_context.output['probabilities'] = softmax_planar(out['postnet_logits'][0])


class AuxiliarySegmPredictStep(_BatchStepBase):
    """The auxiliary-segmentation method (bin-dl/brats_test_auxiliary_segm.py:48-69): the net sees the images plus the
    existing prediction (`labels[:, 1]`) as one more input channel."""

    def __call__(self, batch_context, task_context, context) -> None:
        _check_context(context)
        batch_context.input['images'] = batch_context.input['images'].float().to(context.device, non_blocking=True)
        batch_context.input['labels'] = batch_context.input['labels'].long().to(context.device)
        pred = batch_context.input['labels'][:, 1]
        inpt = torch.cat([batch_context.input['images'], pred.unsqueeze(1).float()], dim=1)
        engine = engine_for(context.model, context.device)
        logits = engine.forward_samples(inpt, 1, dropout_mode=0)[0]
        batch_context.output['logits'] = logits.permute(0, 3, 1, 2)
        batch_context.output['probabilities'] = softmax_planar(logits)
        batch_context.output['orig_prediction'] = pred.unsqueeze(1)


class McPredictStep(_BatchStepBase):
    """T stochastic forwards + the deterministic weight-scaling forward, as ONE folded batch of (T+1)·N images."""

    def __init__(self, mc_steps, defer_ws=True) -> None:
        super().__init__()
        self.mc_steps = mc_steps
        self.defer_ws = defer_ws
        self.slices_seen = 0  # run-global slice index: the Philox stream does not depend on batching

    def __call__(self, batch_context, task_context, context) -> None:
        _check_context(context)
        batch_context.input['images'] = batch_context.input['images'].float().to(context.device, non_blocking=True)
        images = batch_context.input['images']
        engine = engine_for(context.model, context.device, _context_seed(context))
        mode = 1 if engine.dropout else 0
        # the step only ever takes softmaxes of the class logits (customsteps.py:25,33), and a two-class softmax is a function
        # of l0 - l1 alone: the head writes that difference (4 instead of 8 bytes per voxel-sample, bit-identical probabilities)
        diff = LOGIT_DIFF
        logits = engine.forward_samples(images, self.mc_steps + 1, dropout_mode=mode, det_first=True,
                                        slice_index0=self.slices_seen, sample0=0, diff=diff)
        self.slices_seen += images.shape[0]
        # the weight-scaling softmax rides in the fused summary's launch when one follows (it does in every reference
        # script, bin-dl/brats_test_default.py:46-48); anything else that touches the entry first computes it on the spot
        batch_context.output['ws_probabilities'] = LazyWsProbabilities(logits[0], diff) if self.defer_ws else softmax_planar(logits[0], diff)
        batch_context.output['multi_probabilities'] = LazyMultiProbabilities(logits[1:], diff)


class EnsemblePredictionStep(_BatchStepBase):

    def __init__(self, additional_models) -> None:
        super().__init__()
        self.additional_models = additional_models

    def __call__(self, batch_context, task_context, context) -> None:
        _check_context(context)
        batch_context.input['images'] = batch_context.input['images'].float().to(context.device, non_blocking=True)
        images = batch_context.input['images']
        members = [context.model] + list(self.additional_models)
        n, _, h, w = images.shape
        logits = torch.empty((len(members), n, h, w, 2), dtype=torch.float32, device=images.device)
        for m, member in enumerate(members):
            logits[m] = engine_for(member, context.device).forward_samples(images, 1, dropout_mode=0)[0]
        batch_context.output['multi_probabilities'] = LazyMultiProbabilities(logits)


class MultiPredictionSummary(_BatchStepBase):

    def __init__(self, do_mi=False, do_var=False, remove_multi_probs=True, emit_prediction=False, emit_foreground=False) -> None:
        super().__init__()
        self.do_mi = do_mi
        self.do_var = do_var
        self.remove_multi_probs = remove_multi_probs
        self.emit_prediction = emit_prediction  # extra 'prediction' (N, 1, H, W) uint8 output for the in-memory metric path
        self.emit_foreground = emit_foreground  # extra 'foreground' (N, 1, H, W) float32 = probabilities[:, 1], dense

    def __call__(self, batch_context, task_context, context) -> None:
        if self.remove_multi_probs:
            multi = batch_context.output.pop('multi_probabilities')
        else:
            multi = batch_context.output['multi_probabilities']
        ws = batch_context.output.get('ws_probabilities')
        ws = ws if (isinstance(ws, LazyWsProbabilities) and ws._tensor is None and isinstance(multi, LazyMultiProbabilities)
                    and ws.diff == multi.diff) else None
        out = summarize(multi, self.do_mi, self.do_var, self.emit_prediction, self.emit_foreground, ws_logits=None if ws is None else ws.logits)
        if ws is not None:
            ws._tensor = out.pop('ws_probabilities')
        if isinstance(batch_context.output.get('ws_probabilities'), LazyWsProbabilities):
            batch_context.output['ws_probabilities'] = batch_context.output['ws_probabilities'].materialize()
        if not self.remove_multi_probs and isinstance(multi, LazyMultiProbabilities):
            batch_context.output['multi_probabilities'] = multi.materialize()
        # like every reference output the extra maps carry a channel dimension, (N, 1, H, W): the unmodified test loop
        # applies th.channel_to_end to whatever it assembles (common/trainloop/loops.py:210-216)
        for key in ('prediction', 'foreground'):
            if key in out:
                out[key] = out[key].unsqueeze(1)
        batch_context.output.update(out)


def summarize(multi, do_mi=False, do_var=False, emit_prediction=False, emit_foreground=False, ws_logits=None):
    """mean / entropy / [mutual_info] / [variance] / [prediction] / [foreground = mean[:, 1], dense] of a LazyMultiProbabilities or a real
    (T, N, 2, H, W) probability tensor — one fused pass (rechun/dl/customsteps.py:57-71)."""
    if isinstance(multi, LazyMultiProbabilities):
        src, kind = multi.logits, (3 if multi.diff else 0)
        t, n, h, w = src.shape[:4]
    else:
        if multi.dim() != 5 or multi.shape[2] != 2:
            raise ValueError('multi_probabilities must have shape (T, N, 2, H, W), got {}'.format(tuple(multi.shape)))
        src, kind = multi.float().contiguous(), 1
        t, n, c, h, w = src.shape
        if not src.is_cuda:
            raise _lib.RcuError('multi_probabilities must live on the GPU (there is no CPU fallback)')
    dev = src.device
    ws_shape = (n, h, w) if kind == 3 else (n, h, w, 2)
    ext = _torch_ext.ops() if kind in (0, 3) else None
    if ext is not None:
        # torch-extension binding: one registered operator (outputs allocated inside, PyTorch's current stream)
        if ws_logits is not None and tuple(ws_logits.shape) != ws_shape:
            raise ValueError('ws_logits must have shape {} (the layout of the samples)'.format(ws_shape))
        op = ext.aggregate_diff if kind == 3 else ext.aggregate
        mean, entropy, mi, var, pred, fg, ws_out = op(src.contiguous(), None if ws_logits is None else ws_logits.contiguous(), bool(do_mi),
                                                      bool(do_var), bool(emit_prediction), bool(emit_foreground))
        out = {'probabilities': mean, 'entropy': entropy}
        if ws_logits is not None:
            out['ws_probabilities'] = ws_out
        if do_mi:
            out['mutual_info'] = mi
        if do_var:
            out['variance'] = var
        if emit_prediction:
            out['prediction'] = pred
        if emit_foreground:
            out['foreground'] = fg
        return out
    mean = torch.empty((n, 2, h, w), dtype=torch.float32, device=dev)
    entropy = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev)
    mi = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev) if do_mi else None
    var = torch.empty((n, 1, h, w), dtype=torch.float32, device=dev) if do_var else None
    pred = torch.empty((n, h, w), dtype=torch.uint8, device=dev) if emit_prediction else None
    fg = torch.empty((n, h, w), dtype=torch.float32, device=dev) if emit_foreground else None
    ws_out = None
    with torch.cuda.device(dev):
        if ws_logits is not None:
            # interleaved logits (N, H, W, 2) of the deterministic weight-scaling pass: its softmax shares this launch
            if kind not in (0, 3) or tuple(ws_logits.shape) != ws_shape:
                raise ValueError('ws_logits must have shape {} (the layout of the samples)'.format(ws_shape))
            ws_out = torch.empty((n, 2, h, w), dtype=torch.float32, device=dev)
            fn = _lib.lib().rcu_aggregate_ws_diff if kind == 3 else _lib.lib().rcu_aggregate_ws
            _lib.check(fn(_lib.ptr(src.contiguous()), t, n, h * w, _lib.ptr(ws_logits.contiguous()), _lib.ptr(ws_out), _lib.ptr(mean),
                          _lib.ptr(entropy), _lib.ptr(mi), _lib.ptr(var), _lib.ptr(pred), _lib.ptr(fg), _lib.current_stream()))
        else:
            _lib.check(_lib.lib().rcu_aggregate(_lib.ptr(src), kind, t, n, h * w, _lib.ptr(mean), _lib.ptr(entropy), _lib.ptr(mi),
                                                _lib.ptr(var), _lib.ptr(pred), _lib.ptr(fg), None, _lib.current_stream()))
    out = {'probabilities': mean, 'entropy': entropy}
    if ws_out is not None:
        out['ws_probabilities'] = ws_out
    if do_mi:
        out['mutual_info'] = mi
    if do_var:
        out['variance'] = var
    if emit_prediction:
        out['prediction'] = pred
    if emit_foreground:
        out['foreground'] = fg
    return out


class EvalSubjectStep:
    """Drop-in for the subject step of the test scripts (common/trainloop/steps.py:117-131,
    bin-dl/brats_test_default.py:63-77): prediction = argmax of the assembled probabilities, Dice against the
    subject's labels.  Works on what assembly.DeviceSubjectAssembler produces (CUDA tensors: argmax and the confusion
    counts run on the device, only four integers come back) as well as on the numpy arrays of the stock assembler."""

    def __init__(self) -> None:
        from . import evaluation
        self.evaluate = evaluation.ComposeEvaluation([evaluation.DiceNumpy()])

    def __call__(self, subject_context, task_context, context) -> None:
        probabilities = subject_context.subject_data['probabilities']
        if torch.is_tensor(probabilities):
            prediction = (probabilities[..., 1] > probabilities[..., 0]).to(torch.uint8)  # np.argmax: ties -> class 0
        else:
            import numpy as np
            prediction = np.argmax(probabilities, axis=-1)
        to_eval = {'prediction': prediction, 'probabilities': probabilities, 'target': subject_context.subject_data['labels']}
        results = {}
        self.evaluate(to_eval, results)
        subject_context.metrics.update(results)


def _context_seed(context):
    try:
        seed = context.get_seed()
    except Exception:
        seed = None
    return seed
