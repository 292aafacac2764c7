#!/usr/bin/env python
"""Stage the UNMODIFIED reference under oracle/_ref/ (git-ignored build output; travels to the GPU box with
the working tree) -- TEST / BASELINE INFRASTRUCTURE ONLY, never imported by the product package.

    python oracle/stage_ref.py            # copy + prebuild the reference's JIT extensions for sm_100a

What it is for:
  * `bench.py --impl reference` and `cpu_baseline` time the reference's OWN `model.py` + `op/` (native CPU
    fallbacks op/fused_act.py:105-116, op/upfirdn2d.py:146-200) on the host cores;
  * the `reference_gpu` leg times the reference CUDA path (its SIMT `op/*.cu` compiled for sm_100a + cuDNN
    grouped convolutions) on the same B200 -- the real competitor;
  * the compat launcher (`python -m b200gan.run oracle/_ref/prune.py ...`) runs the unmodified driver scripts
    against the drop-in `model` / `op`.

Nothing is edited: files are byte-for-byte copies (sha256 manifest in oracle/_ref/MANIFEST.json).  The JIT
extensions (`torch.utils.cpp_extension.load`, op/fused_act.py:11-17, op/upfirdn2d.py:10-16) are pre-built into
oracle/_ref/_ext with TORCH_CUDA_ARCH_LIST=10.0a so that the GPU box reuses them (it has no need to wait ~100 s).
/root/reference does not exist on the GPU box; nothing reads it at run time.
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get('CAGC_REFERENCE', '/root/reference')
DST = os.path.join(HERE, '_ref')
EXT = os.path.join(DST, '_ext')

# python sources of the path + its drivers, the two native ops, the small vendored weights.  Left out: doc/,
# the 66 MB FFHQ Inception statistics (get_fid.py's FID arithmetic is outside the hot path; the launcher feeds
# synthetic statistics), Miscellaneous/ (dead or TF-dependent scripts).
INCLUDE = ['model.py', 'train.py', 'train_hyperparams.py', 'prune.py', 'get_fid.py', 'get_ppl.py', 'dataset.py',
           'op', 'Util', 'lpips', 'Evaluation/fid.py', 'Evaluation/ppl.py', 'Evaluation/inception.py',
           'Evaluation/calc_inception.py', 'Miscellaneous/distributed.py', 'LICENSE', 'LICENSE-NVIDIA', 'LICENSE-LPIPS',
           'LICENSE-FID']
SKIP_DIRS = {'__pycache__', 'modules'}          # Util/face_parsing/modules = unused in-place ABN
# 53 MB of parser weights: only needed to run prune.py / train.py with the real face parser (the tests use the
# synthetic mask); left out by default because every gpurun call pushes the tree (--with-bisenet-weights adds them)
BISENET = 'Util/face_parsing/pretrained_model/79999_iter.pth'


def _sha(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        for blk in iter(lambda: f.read(1 << 20), b''):
            h.update(blk)
    return h.hexdigest()


def env_for_ref():
    env = dict(os.environ)
    env['TORCH_EXTENSIONS_DIR'] = EXT
    env['TORCH_CUDA_ARCH_LIST'] = '10.0a'
    env.setdefault('MAX_JOBS', str(os.cpu_count() or 4))
    return env


def stage(with_bisenet_weights=False, prebuild=True, verbose=True):
    if not os.path.isdir(REF_SRC):
        if os.path.isdir(DST):
            return DST                     # GPU box: use what travelled with the tree
        raise RuntimeError(f'{REF_SRC} not found and nothing staged under {DST}')
    manifest = {}
    for item in INCLUDE:
        src = os.path.join(REF_SRC, item)
        if os.path.isfile(src):
            files = [item]
        else:
            files = []
            for root, dirs, names in os.walk(src):
                dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
                for n in names:
                    files.append(os.path.relpath(os.path.join(root, n), REF_SRC))
        for rel in files:
            if rel == BISENET and not with_bisenet_weights:
                continue
            s, d = os.path.join(REF_SRC, rel), os.path.join(DST, rel)
            digest = _sha(s)
            manifest[rel] = digest
            if os.path.exists(d) and _sha(d) == digest:
                continue
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': REF_SRC, 'files': manifest}, f, indent=1, sort_keys=True)
    if prebuild:
        built = all(os.path.exists(os.path.join(EXT, n, n + '.so')) for n in ('fused', 'upfirdn2d'))
        if not built:
            if verbose:
                print('stage_ref: building the reference JIT extensions for sm_100a (about 100 s) ...', flush=True)
            os.makedirs(EXT, exist_ok=True)
            subprocess.run([sys.executable, '-c', 'import model'], cwd=DST, env=env_for_ref(), check=True)
    if verbose:
        n = len(manifest)
        size = sum(os.path.getsize(os.path.join(DST, r)) for r in manifest) / 1e6
        print(f'stage_ref: {n} files ({size:.1f} MB) under {DST}')
    return DST


if __name__ == '__main__':
    stage(with_bisenet_weights='--with-bisenet-weights' in sys.argv, prebuild='--no-build' not in sys.argv)
