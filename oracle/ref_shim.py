"""TEST INFRASTRUCTURE ONLY — loads the *unmodified* reference from /root/reference.

The reference (alainjungo/reliability-challenges-uncertainty) is pure Python but
imports third-party packages that are absent here (pymia==0.2.1, tensorboardX,
SimpleITK).  This module registers minimal stand-ins in ``sys.modules`` so that
the reference's own modules on the hot path import and run on CPU:

  common.model.unet, common.utils.torchhelper, common.trainloop.{context,steps},
  rechun.dl.customsteps, common.evalutation.{numpyfunctions,eval},
  rechun.eval.{analysis,helper}

Only ``tests/`` (golden-vector generation, run in the authoring container where
/root/reference exists) may import this.  It cannot travel to the GPU box;
``oracle/restate.py`` is the travelling restatement, pinned against fixtures
that ``tests/golden/make_golden.py`` generates through this module.

pymia.evaluation.metric is the one piece of *arithmetic* living in an absent
dependency (pymia 0.2.1, requirements.txt:5).  Its three classes used at
common/evalutation/numpyfunctions.py:128-151 are restated below from pymia's
published semantics (ConfusionMatrix: tp/tn/fp/fn by ==1/==0 comparisons,
n = prediction.size; DiceCoefficient = 2tp/(2tp+fp+fn), 1.0 when everything is
empty; Accuracy = (tp+tn)/(tp+tn+fp+fn)).  They cannot be diffed against the
pinned wheel offline: that sub-part of parity is "unpinned" (see DESIGN.md).
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get('RCU_REFERENCE_ROOT', '/root/reference')


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'common', 'model'))


class _Anything:
    """Attribute-permissive dummy used for pymia classes that are only subclassed/named."""

    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return None


def _permissive_module(name):
    mod = types.ModuleType(name)

    def _getattr(attr):
        if attr.startswith('__'):
            raise AttributeError(attr)
        cls = type(attr, (_Anything,), {})
        setattr(mod, attr, cls)
        return cls

    mod.__getattr__ = _getattr
    return mod


# --- pymia.evaluation.metric (restated, see module docstring) -----------------------------------
class ConfusionMatrix:
    def __init__(self, prediction: np.ndarray, label: np.ndarray):
        self.tp = np.sum(np.logical_and(prediction == 1, label == 1))
        self.tn = np.sum(np.logical_and(prediction == 0, label == 0))
        self.fp = np.sum(np.logical_and(prediction == 1, label == 0))
        self.fn = np.sum(np.logical_and(prediction == 0, label == 1))
        self.n = prediction.size


class _ConfusionMatrixMetric:
    def __init__(self):
        self.confusion_matrix = None


class DiceCoefficient(_ConfusionMatrixMetric):
    def calculate(self):
        cm = self.confusion_matrix
        if (cm.tp == 0) and ((cm.tp + cm.fp + cm.fn) == 0):
            return 1.
        return 2 * cm.tp / (2 * cm.tp + cm.fp + cm.fn)


class Accuracy(_ConfusionMatrixMetric):
    def calculate(self):
        cm = self.confusion_matrix
        sum_ = cm.tp + cm.tn + cm.fp + cm.fn
        if sum_ != 0:
            return (cm.tp + cm.tn) / sum_
        return 0


def _install_shims():
    if 'pymia' in sys.modules and getattr(sys.modules['pymia'], '_rcu_shim', False):
        return
    pymia = types.ModuleType('pymia')
    pymia._rcu_shim = True
    pymia.__path__ = []

    config_pkg = types.ModuleType('pymia.config')
    config_pkg.__path__ = []
    configuration = types.ModuleType('pymia.config.configuration')

    class Dictable:
        def to_dict(self, **kwargs):
            return dict(vars(self))

        def from_dict(self, d, **kwargs):
            for k, v in d.items():
                setattr(self, k, v)

    class ConfigurationBase(Dictable):
        @classmethod
        def version(cls):
            return 1

        @classmethod
        def type(cls):
            return cls.__name__

    def member_to_dict(target_dict, source_object):
        target_dict.update(vars(source_object))

    def dict_to_member(target_object, source_dict):
        for k, v in source_dict.items():
            setattr(target_object, k, v)

    def load(file_path, config_cls):
        raise NotImplementedError('config loading is outside the oracle scope')

    def save(file_path, config):
        raise NotImplementedError('config saving is outside the oracle scope')

    configuration.Dictable = Dictable
    configuration.ConfigurationBase = ConfigurationBase
    configuration.member_to_dict = member_to_dict
    configuration.dict_to_member = dict_to_member
    configuration.load = load
    configuration.save = save
    configuration.__all__ = ['Dictable', 'ConfigurationBase', 'member_to_dict', 'dict_to_member', 'load', 'save']

    data_pkg = _permissive_module('pymia.data')
    data_pkg.__path__ = []
    subs = {}
    for sub in ('extraction', 'transformation', 'assembler', 'conversion', 'subjectfile', 'creation',
                'indexexpression'):
        subs[sub] = _permissive_module('pymia.data.' + sub)
        setattr(data_pkg, sub, subs[sub])

    evaluation_pkg = types.ModuleType('pymia.evaluation')
    evaluation_pkg.__path__ = []
    metric = types.ModuleType('pymia.evaluation.metric')
    metric.ConfusionMatrix = ConfusionMatrix
    metric.DiceCoefficient = DiceCoefficient
    metric.Accuracy = Accuracy
    evaluation_pkg.metric = metric

    pymia.config = config_pkg
    config_pkg.configuration = configuration
    pymia.data = data_pkg
    pymia.evaluation = evaluation_pkg

    sys.modules['pymia'] = pymia
    sys.modules['pymia.config'] = config_pkg
    sys.modules['pymia.config.configuration'] = configuration
    sys.modules['pymia.data'] = data_pkg
    for sub, mod in subs.items():
        sys.modules['pymia.data.' + sub] = mod
    sys.modules['pymia.evaluation'] = evaluation_pkg
    sys.modules['pymia.evaluation.metric'] = metric

    if 'tensorboardX' not in sys.modules:
        tbx = types.ModuleType('tensorboardX')
        tbx.SummaryWriter = type('SummaryWriter', (_Anything,), {})
        sys.modules['tensorboardX'] = tbx
    if 'SimpleITK' not in sys.modules:
        sys.modules['SimpleITK'] = _permissive_module('SimpleITK')
    try:  # rechun/eval/hook.py:4-5 imports matplotlib for its PDF hooks; the CSV hooks next to them do not use it
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = _permissive_module('matplotlib')
        mpl.__path__ = []
        backends = _permissive_module('matplotlib.backends')
        backends.__path__ = []
        backends.backend_pdf = _permissive_module('matplotlib.backends.backend_pdf')
        mpl.backends, mpl.pyplot = backends, _permissive_module('matplotlib.pyplot')
        sys.modules.update({'matplotlib': mpl, 'matplotlib.backends': backends,
                            'matplotlib.backends.backend_pdf': backends.backend_pdf, 'matplotlib.pyplot': mpl.pyplot})
    if 'h5py' not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            sys.modules['h5py'] = _permissive_module('h5py')


def load():
    """Make ``import common...`` / ``import rechun...`` resolve to the reference tree."""
    if not available():
        raise RuntimeError('reference tree not found at {} (it does not exist on the GPU box)'.format(REFERENCE_ROOT))
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # np.bool is used at common/evalutation/eval.py:159-160,184-185; numpy>=2 provides it again, numpy 1.24-1.26 not.
    if not hasattr(np, 'bool'):
        np.bool = bool  # pragma: no cover
