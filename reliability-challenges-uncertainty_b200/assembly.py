"""Device-side subject assembly.

The reference's test loop hands every selected batch output to a pymia assembler after `channel_to_end` and
`tensor_to_numpy` (common/trainloop/loops.py:210-223): one device->host sync per batch and one Python-level copy per
slice.  The two classes below follow the same assembler protocol the loop relies on —

    add_batch(to_assemble, batch, last_batch=False)      loops.py:223
    subjects_ready                                       loops.py:225-226
    get_assembled_subject(subject_index)                 loops.py:227

— but keep the subject volumes as device tensors, so `loop.Test(steps, subject_steps, DeviceSubjectAssembler(),
entries=..., convert_fn=None)` runs the unmodified loop without leaving HBM.  `SubjectAssembler` /
`Subject2dAssembler` themselves live in pymia 0.2.1 (requirements.txt:5), whose source is not part of the reference
tree: their bookkeeping (a subject is ready when a sample of the next subject arrives or on the last batch; batch
keys 'subject_index', 'index_expr', 'shape'; pickled index expressions) is restated here from the call sites
bin-dl/brats_test_default.py:53-54 and bin-dl/isic_test_default.py:54-55.

Consecutive slices of one subject are written with ONE strided copy per run and entry (the permuted NHWC view of the
NCHW step output is read in place), not one copy per slice.
"""
import pickle

import torch


def _expression_of(index_expr):
    if isinstance(index_expr, (bytes, bytearray)):
        index_expr = pickle.loads(index_expr)
    expression = getattr(index_expr, 'expression', index_expr)
    if isinstance(expression, list):
        expression = tuple(expression)
    if not isinstance(expression, tuple):
        expression = (expression,)
    return expression


def _leading_int(expression):
    """k when the expression selects slice k along axis 0 and everything else fully, else None."""
    if len(expression) == 0 or isinstance(expression[0], bool) or not isinstance(expression[0], int):
        return None
    for e in expression[1:]:
        if not (isinstance(e, slice) and e == slice(None)):
            return None
    return int(expression[0])


def _subject_shape(shape, expression, sample):
    shape = tuple(int(s) for s in shape)
    # region the expression selects in a volume of `shape`
    region = torch.empty(shape, device='meta')[expression].shape
    extra = tuple(sample.shape[len(region):]) if sample.dim() > len(region) else ()
    if tuple(sample.shape[:len(region)]) != tuple(region):
        raise ValueError('sample of shape {} does not fit index expression {} of a subject of shape {}'.format(
            tuple(sample.shape), expression, shape))
    return shape + extra


class DeviceSubjectAssembler:
    """Assembles slice-wise (or patch-wise) batch outputs into per-subject device tensors."""

    def __init__(self, device=None) -> None:
        self.device = device
        self.predictions = {}
        self.subjects_ready = set()

    def _check_batch(self, batch):
        for key, extractor in (('subject_index', 'IndexingExtractor'), ('index_expr', 'IndexingExtractor'),
                               ('shape', 'ImageShapeExtractor')):
            if key not in batch:
                raise ValueError('DeviceSubjectAssembler requires "{}" to be extracted (use {})'.format(key, extractor))

    def add_batch(self, to_assemble, batch: dict, last_batch=False):
        self._check_batch(batch)
        if not isinstance(to_assemble, dict):
            to_assemble = {'__prediction': to_assemble}
        subject_indices = [int(s) for s in batch['subject_index']]
        expressions = [_expression_of(e) for e in batch['index_expr']]
        n = len(subject_indices)
        i = 0
        while i < n:
            subject = subject_indices[i]
            if subject not in self.predictions:
                if self.predictions:  # a new subject starts: everything assembled so far is complete
                    self.subjects_ready = set(self.predictions.keys())
                self.predictions[subject] = self._init_new_subject(to_assemble, batch, i, expressions[i])
            # longest run of consecutive axis-0 slices of this subject
            k0 = _leading_int(expressions[i])
            j = i + 1
            if k0 is not None:
                while j < n and subject_indices[j] == subject and _leading_int(expressions[j]) == k0 + (j - i):
                    j += 1
            for key, value in to_assemble.items():
                dst = self.predictions[subject][key]
                if k0 is not None:
                    dst[k0:k0 + (j - i)].copy_(value[i:j], non_blocking=True)
                else:
                    dst[expressions[i]] = value[i]
            i = j
        if last_batch:
            self.end()

    def end(self):
        self.subjects_ready = set(self.predictions.keys())

    def _init_new_subject(self, to_assemble, batch, idx, expression):
        subject = {}
        for key, value in to_assemble.items():
            if not torch.is_tensor(value):
                raise ValueError('entry "{}" is not a tensor (use convert_fn=None with the device assembler)'.format(key))
            shape = _subject_shape(batch['shape'][idx], expression, value[idx])
            subject[key] = torch.zeros(shape, dtype=value.dtype, device=self.device if self.device is not None else value.device)
        return subject

    def get_assembled_subject(self, subject_index: int):
        try:
            self.subjects_ready.remove(subject_index)
        except KeyError:
            if subject_index not in self.predictions:
                raise ValueError('Subject with index {} not in assembler'.format(subject_index))
        assembled = self.predictions.pop(subject_index)
        if '__prediction' in assembled:
            return assembled['__prediction']
        return assembled


class DeviceSubject2dAssembler:
    """One sample = one subject (ISIC images, bin-dl/isic_test_default.py:54): entries stay device tensors."""

    def __init__(self) -> None:
        self.predictions = {}
        self.subjects_ready = set()

    def add_batch(self, to_assemble, batch: dict, last_batch=False):
        if 'subject_index' not in batch:
            raise ValueError('DeviceSubject2dAssembler requires "subject_index" to be extracted (use IndexingExtractor)')
        if not isinstance(to_assemble, dict):
            to_assemble = {'__prediction': to_assemble}
        for idx, subject in enumerate(batch['subject_index']):
            subject = int(subject)
            entry = self.predictions.setdefault(subject, {})
            for key, value in to_assemble.items():
                entry[key] = value[idx]
            self.subjects_ready.add(subject)

    def get_assembled_subject(self, subject_index: int):
        try:
            self.subjects_ready.remove(subject_index)
        except KeyError:
            if subject_index not in self.predictions:
                raise ValueError('Subject with index {} not in assembler'.format(subject_index))
        assembled = self.predictions.pop(subject_index)
        if '__prediction' in assembled:
            return assembled['__prediction']
        return assembled
