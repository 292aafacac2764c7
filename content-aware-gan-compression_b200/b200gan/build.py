"""Ahead-of-time build of lib/libcagc_b200.so (no JIT at import, unlike op/fused_act.py:11-17).

nvcc cross-compiles for sm_100a without a GPU; the resulting .so is git-ignored and travels to the
GPU box with the working tree.
"""
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
LIB = os.path.join(PKG, 'lib', 'libcagc_b200.so')
STAMP = LIB + '.stamp'
SOURCES = ['ops.cu', 'conv_simt.cu', 'nhwc_aux.cu', 'conv_tc.cu', 'prep.cu', 'disc.cu', 'linear.cu', 'lpips.cu', 'kdloss.cu']
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '-shared']


def _fingerprint():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ['../../include/cagc_b200.h']:
        path = os.path.normpath(os.path.join(CSRC, name))
        if os.path.isfile(path):
            h.update(name.encode())
            with open(path, 'rb') as f:
                h.update(f.read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    fp = _fingerprint()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == fp:
        return LIB
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolkit: use the prebuilt library that travelled with the tree
        raise RuntimeError('nvcc not found and no prebuilt libcagc_b200.so')
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [nvcc] + FLAGS + ['-I', os.path.join(ROOT, 'include'), '-I', CSRC] + \
          [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB]
    if verbose:
        cmd.insert(1, '-Xptxas=-v')
        print(' '.join(cmd))
    subprocess.run(cmd, check=True)
    with open(STAMP, 'w') as f:
        f.write(fp)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
