import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run on the GPU box with -m gpu)')


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='no CUDA device in this container (GPU tests run under gpurun)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """The C-ABI library must exist for every suite (CPU tests load it and check its symbols)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('rcu_build', os.path.join(ROOT, 'reliability-challenges-uncertainty_b200', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    yield


@pytest.fixture(scope='session')
def golden_unet():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'unet_golden.npz'))


@pytest.fixture(scope='session')
def golden_metrics():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'metrics_golden.npz'), allow_pickle=False)


@pytest.fixture(scope='session')
def golden_aux():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'aux_golden.npz'), allow_pickle=False)


@pytest.fixture(scope='session')
def golden_tables():
    import numpy as np
    return np.load(os.path.join(ROOT, 'tests', 'golden', 'tables_golden.npz'), allow_pickle=False)
