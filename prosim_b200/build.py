"""Build libprosim_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'api.cu')
OUT = os.path.join(HERE, 'libprosim_b200.so')
DEPS = [os.path.join(HERE, 'csrc', f) for f in os.listdir(os.path.join(HERE, 'csrc'))] + \
    [os.path.join(os.path.dirname(HERE), 'include', 'prosim_b200.h')]


def nvcc_cmd():
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    return [nvcc, '-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo',
            '-Xcompiler', '-fPIC', '-shared', '-o', OUT, SRC]


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = nvcc_cmd()
    if verbose:
        print(' '.join(cmd))
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
