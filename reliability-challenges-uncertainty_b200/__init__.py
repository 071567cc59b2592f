"""rcu_b200 — B200-native (sm_100a) stochastic-inference + uncertainty hot path of
alainjungo/reliability-challenges-uncertainty, behind the reference's own Python protocols.

  model.B200UNet                      `context.model` drop-in (tcgen05 implicit-GEMM U-Net forward)
  steps.{SegmentationPredictStep, McPredictStep, EnsemblePredictionStep, MultiPredictionSummary}
  evaluation.{EceBinaryNumpy, UncertaintyErrorDiceNumpy, UncertaintyAndCorrectionEvalNumpy, DiceNumpy,
              ConfusionMatrix, ComposeEvaluation, ...} and the np_fn twins (ece_binary, uncertainty, dice, ...)
  metrics / tables                    batched device tables and the host logic around them
  hooks.DeviceMetricsHook             TestLoopHook that evaluates subjects on the device
  assembly.DeviceSubjectAssembler     subject assembly of the test loop on device tensors (no per-batch D2H)
  nifti.AsyncNiftiWriteHook           optional {subject}_probabilities / _prediction .nii.gz files, written in the background
  distributed                         slice / sample / member sharding over one process per GPU

The directory name follows the build contract (`reliability-challenges-uncertainty_b200/`); import it as
`rcu_b200` (the repo-root shim `rcu_b200.py` registers this directory under that name).
"""
from . import _lib  # noqa: F401
from . import tables  # noqa: F401

__all__ = ['_lib', 'tables', 'metrics', 'evaluation', 'model', 'steps', 'hooks', 'distributed', 'assembly', 'synth', 'nifti']


def __getattr__(name):
    # torch-dependent submodules are imported on first use so that `import rcu_b200` stays cheap
    if name in ('metrics', 'evaluation', 'model', 'steps', 'hooks', 'distributed', 'assembly', 'synth', 'nifti'):
        import importlib
        return importlib.import_module('.' + name, __name__)
    raise AttributeError(name)
