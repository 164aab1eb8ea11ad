"""Builds librlgs.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

    python -m rlgpuschedule_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc')
SO = os.path.join(HERE, 'librlgs.so')
ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']


def nvcc_path():
    p = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(p):
        raise RuntimeError('nvcc not found')
    return p


def sources():
    return sorted(os.path.join(SRC, f) for f in os.listdir(SRC) if f.endswith('.cu'))


def deps():
    out = [os.path.join(SRC, f) for f in os.listdir(SRC)]
    out.append(os.path.join(os.path.dirname(HERE), 'include', 'rlgs.h'))
    return out


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in deps())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [nvcc_path()] + ARCH + ['-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '-shared',
                                   '--fmad=false', '-Xptxas', '-v' if verbose else '-O3',
                                   '-o', SO] + sources()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + r.stdout)
    if verbose:
        print(r.stdout)
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
