"""The torch-extension binding of the C-ABI: `torch.ops.rcu_b200.*` custom operators built from csrc/torch_binding.cpp
(`librcu_b200_torch.so`, next to `librcu_b200.so`; build.py / __graft_entry__.build() compile both).

The ctypes binding in _lib.py remains the reference binding of every entry point; this module routes the per-call hot
entries (the fused metric pass, the aggregation, the U-Net forward) through registered torch operators instead: they take
tensors, launch on PyTorch's current stream and skip the ctypes argument marshalling.  `RCU_B200_BINDING=ctypes` switches
the routing off (A/B; tools/binding_overhead.py measures both).  There is no fallback in the other direction either: if
the extension is missing the ctypes path is the one that runs, and without librcu_b200.so nothing runs.
"""
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
EXT_PATH = os.path.join(_HERE, 'librcu_b200_torch.so')
_state = {'ops': None, 'tried': False}


def ops():
    """torch.ops.rcu_b200 (loading the extension on first use), or None when it is not built / switched off."""
    if not _state['tried']:
        _state['tried'] = True
        if os.environ.get('RCU_B200_BINDING', 'torch') != 'ctypes' and os.path.exists(EXT_PATH):
            from . import _lib
            _lib.lib()                                  # librcu_b200.so first: the extension resolves its symbols against it
            torch.ops.load_library(EXT_PATH)
            if int(torch.ops.rcu_b200.abi_version()) != _lib.RCU_ABI_VERSION:
                raise _lib.RcuError('librcu_b200_torch.so was built against another ABI version (rebuild: python __graft_entry__.py)')
            _state['ops'] = torch.ops.rcu_b200
    return _state['ops']
