"""The C-ABI library loads on a CPU-only box, exports every symbol include/rcu_b200.h declares, validates its
arguments and fails loudly (never silently falls back) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import rcu_b200  # noqa: F401  (registers the package directory under this name)
from rcu_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'rcu_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(rcu_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_are_exported_and_bound():
    declared = _declared_symbols()
    assert len(declared) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), 'library does not export {}'.format(name)
    assert sorted(_lib.PROTOTYPES) == declared, 'ctypes prototypes drifted from the header'


def test_abi_version_and_error_channel():
    lib = _lib.lib()
    assert lib.rcu_abi_version() == _lib.RCU_ABI_VERSION == 3
    rc = lib.rcu_metrics_workspace_init(None, 0, None)
    assert rc == _lib.RCU_EINVAL and 'NULL' in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    assert lib.rcu_metrics_workspace_bytes(1) >= 2048 * 8
    assert lib.rcu_metrics_workspace_bytes(5000) > lib.rcu_metrics_workspace_bytes(1)


def test_argument_validation_without_gpu():
    lib = _lib.lib()
    edges = (ctypes.c_float * 11)(*np.linspace(0, 1, 11))
    one = ctypes.c_void_p(16)  # never dereferenced: validation fails first
    assert lib.rcu_calib_hist(None, one, None, 10, 1, edges, 10, 0.0, 1.0, one, one, one, one, 1 << 22, None) == _lib.RCU_EINVAL
    assert lib.rcu_calib_hist(one, one, None, 10, 1, edges, 64, 0.0, 1.0, one, one, one, one, 1 << 22, None) == _lib.RCU_EINVAL
    assert 'n_bins' in _lib.last_error()
    bad_edges = (ctypes.c_float * 11)(*([0.5] * 11))
    assert lib.rcu_calib_hist(one, one, None, 10, 1, bad_edges, 10, 0.0, 1.0, one, one, one, one, 1 << 22, None) == _lib.RCU_EINVAL
    seg = (ctypes.c_uint8 * 3)(0, 1, 5)
    br = (ctypes.c_float * 2)(0.1, 0.2)
    assert lib.rcu_ue_hist(one, 0, one, one, None, 10, 1, br, None, 2, seg, 3, one, None, one, 1 << 22, None) == _lib.RCU_EINVAL
    assert 'seg_class' in _lib.last_error()
    assert lib.rcu_ue_hist(one, 7, one, one, None, 10, 1, br, None, 2, seg, 3, one, None, one, 1 << 22, None) == _lib.RCU_EINVAL
    assert lib.rcu_aggregate(one, 0, 3, 2, 15, one, None, None, None, None, None, None, None) == _lib.RCU_EINVAL  # odd hw
    assert lib.rcu_aggregate(one, 9, 3, 2, 16, one, None, None, None, None, None, None, None) == _lib.RCU_EINVAL


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_compute_fails_loudly_without_gpu():
    lib = _lib.lib()
    assert lib.rcu_device_check(0) == _lib.RCU_ECUDA and 'no CUDA device' in _lib.last_error()
    from rcu_b200 import metrics
    with pytest.raises(_lib.RcuError):
        metrics.calibration_tables(np.zeros(8, dtype=np.float32), np.zeros(8, dtype=np.uint8))
    from oracle import restate as R
    from rcu_b200 import model
    cfg = R.UNetConfig()
    with pytest.raises(Exception) as e:
        model.B200UNet(R.init_state_dict(cfg, 0), device='cuda:0')
    assert not isinstance(e.value, (AttributeError, TypeError)), e.value


def test_missing_library_is_an_error(monkeypatch):
    monkeypatch.setattr(_lib, '_lib', None)
    monkeypatch.setattr(_lib, 'LIB_PATH', '/nonexistent/librcu_b200.so')
    with pytest.raises(_lib.RcuError, match='no CPU fallback'):
        _lib.lib()


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'reliability-challenges-uncertainty_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
                assert '/root/reference' not in src, f


def test_torch_extension_registers_the_operators_and_validates_like_the_abi():
    """librcu_b200_torch.so (csrc/torch_binding.cpp): torch.ops.rcu_b200.* over the same C-ABI.  On a CPU-only box the
    operators must load, report the ABI version and refuse host tensors with the reference's exception type."""
    from rcu_b200 import _torch_ext
    ops = _torch_ext.ops()
    if ops is None:
        pytest.skip('torch extension not built (python __graft_entry__.py builds it)')
    assert int(ops.abi_version()) == _lib.RCU_ABI_VERSION
    for name in ('eval_fused', 'aggregate', 'unet_forward'):
        assert hasattr(ops, name)
    z8 = torch.zeros(4, dtype=torch.uint8)
    with pytest.raises(ValueError):
        ops.eval_fused(torch.zeros(4), z8, z8, None, torch.linspace(0, 1, 11), torch.zeros(2), torch.zeros(3, dtype=torch.uint8), 1, 3, torch.zeros(8, dtype=torch.uint8))
    with pytest.raises(ValueError):
        ops.aggregate(torch.zeros(2, 1, 4, 4, 2), None, False, False, False, False)
