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


def build_variant(name, defines):
    """A/B build of the library with extra -D switches (e.g. RCU_ARRIVE_PER_THREAD=1): build/variants/librcu_b200_<name>.so.
    Select it at run time with RCU_B200_LIB=<path> RCU_B200_BINDING=ctypes (the torch extension links the default library)."""
    out_dir = os.path.join(HERE, 'build', 'variants')
    os.makedirs(out_dir, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f not in ('--use_fast_math=false', '-v', '-Xptxas')] + ['-D' + d for d in defines]
    objs = []
    for src in SOURCES:
        obj = os.path.join(out_dir, '%s_%s' % (name, src.replace('.cu', '.o')))
        res = subprocess.run([_nvcc()] + flags + ['-c', os.path.join(CSRC, src), '-o', obj], capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError('nvcc failed on {} ({})'.format(src, name))
        objs.append(obj)
    lib = os.path.join(out_dir, 'librcu_b200_%s.so' % name)
    res = subprocess.run([_nvcc(), '-shared', '-o', lib] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-ldl'], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('link failed')
    for o in objs:
        os.remove(o)
    return lib


TORCH_EXT_PATH = os.path.join(HERE, 'librcu_b200_torch.so')
TORCH_EXT_STAMP = os.path.join(HERE, 'build', 'librcu_b200_torch.stamp')


def build_torch_extension(force=False, verbose=False):
    """Compile csrc/torch_binding.cpp (the torch.ops.rcu_b200.* operators over the C-ABI) in-tree with g++ against the
    installed PyTorch and link it to librcu_b200.so next to it.  Returns the path (load it with torch.ops.load_library)."""
    import torch
    from torch.utils import cpp_extension
    src = os.path.join(CSRC, 'torch_binding.cpp')
    h = hashlib.sha256()
    for path in (src, os.path.join(HERE, '..', 'include', 'rcu_b200.h')):
        with open(path, 'rb') as f:
            h.update(f.read())
    h.update(torch.__version__.encode())
    fp = h.hexdigest()
    if not force and os.path.exists(TORCH_EXT_PATH) and os.path.exists(TORCH_EXT_STAMP):
        with open(TORCH_EXT_STAMP) as f:
            if f.read().strip() == fp:
                return TORCH_EXT_PATH
    build(force=False, verbose=verbose)          # the library it links against
    torch_lib = os.path.join(os.path.dirname(torch.__file__), 'lib')
    cuda_inc = os.path.join(os.environ.get('CUDA_HOME', '/usr/local/cuda'), 'include')
    cmd = ['g++', '-O2', '-std=c++17', '-fPIC', '-shared', src, '-o', TORCH_EXT_PATH,
           '-D_GLIBCXX_USE_CXX11_ABI=%d' % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ['-I' + p for p in cpp_extension.include_paths()] + ['-I' + cuda_inc]
    cmd += ['-L' + torch_lib, '-ltorch', '-ltorch_cpu', '-lc10', '-ltorch_cuda', '-lc10_cuda', '-L' + HERE, '-l:librcu_b200.so',
            '-Wl,-rpath,$ORIGIN', '-Wl,-rpath,' + torch_lib, '-Wl,--no-as-needed']
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('g++ failed on torch_binding.cpp')
    os.makedirs(os.path.dirname(TORCH_EXT_STAMP), exist_ok=True)
    with open(TORCH_EXT_STAMP, 'w') as f:
        f.write(fp)
    return TORCH_EXT_PATH


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
    print(build_torch_extension(force='--force' in sys.argv, verbose='-v' in sys.argv))
