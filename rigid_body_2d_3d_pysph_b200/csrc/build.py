"""Build librbx.so in-tree with nvcc for sm_100a (no JIT cache, no torch)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ['rbx_cells.cu', 'rbx_contact.cu', 'rbx_bodies.cu', 'rbx_lvc.cu',
           'rbx_setup.cu', 'rbx_canelas.cu']
LIB = os.path.join(HERE, 'librbx.so')


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES] + \
        [os.path.join(HERE, 'rbx_common.cuh'),
         os.path.join(ROOT, 'include', 'rbx.h')]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: tuning variants (-DRBX_KACC=2 ...) built beside the
    default library; select one at run time with RBX_LIB=<path>."""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc, '-O3', '-std=c++17', '-lineinfo',
           '-gencode', 'arch=compute_100a,code=sm_100a',
           '-Xcompiler', '-fPIC', '-shared',
           '-I', os.path.join(ROOT, 'include'), '-I', HERE]
    if verbose:
        cmd += ['-Xptxas', '-v']
    cmd += ['-D' + d for d in defines]
    cmd += [os.path.join(HERE, s) for s in SOURCES] + ['-o', out or LIB]
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == '__main__':
    build(force=True, verbose='-v' in sys.argv)
    print(LIB)
