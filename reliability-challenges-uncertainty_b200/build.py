"""Builds librcu_b200.so (the C-ABI library declared in include/rcu_b200.h) in-tree with nvcc for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so loads on a CPU-only box (cudart is linked statically and
the driver entry point for TMA descriptor encoding is resolved lazily at run time).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'librcu_b200.so')
STAMP_PATH = os.path.join(HERE, 'build', 'librcu_b200.stamp')
SOURCES = ['error.cu', 'metrics.cu', 'aggregate.cu', 'masks.cu', 'prepare.cu', 'collective.cu', 'unet.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--use_fast_math=false',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; cannot build librcu_b200.so')


def _fingerprint():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ['../../include/rcu_b200.h']
    for name in names:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, 'rb') as f:
                h.update(f.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile if sources changed since the last build. Returns the library path."""
    fp = _fingerprint()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP_PATH):
        with open(STAMP_PATH) as f:
            if f.read().strip() == fp:
                return LIB_PATH
    os.makedirs(os.path.dirname(STAMP_PATH), exist_ok=True)
    objs = []
    flags = [f for f in NVCC_FLAGS if f != '--use_fast_math=false']
    for src in SOURCES:
        obj = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        cmd = [_nvcc()] + flags + ['-c', os.path.join(CSRC, src), '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
        if res.returncode != 0:
            raise RuntimeError('nvcc failed on {}'.format(src))
        with open(obj + '.ptxas.log', 'w') as f:
            f.write(res.stderr)
        objs.append(obj)
    cmd = [_nvcc(), '-shared', '-o', LIB_PATH] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-ldl']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('link failed')
    with open(STAMP_PATH, 'w') as f:
        f.write(fp)
    return LIB_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
