import os
import sys

# tests state their engine explicitly (use_algo / exact_fp32); everything else is checked on the exact-fp32 engine.
# The product default is `auto` = tcgen05 TF32 on sm_100 (b200gan/config.py).
os.environ.setdefault('CAGC_CONV_ALGO', 'simt')

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'content-aware-gan-compression_b200')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')

# The drop-in boundary is the pair of top-level modules `model` and `op`
# (SURVEY.md §8b): put the package dir first on sys.path so `import model`
# binds to ours exactly as the reference scripts would see it.
for p in (PKG, ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
