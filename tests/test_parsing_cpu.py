"""b200gan.parsing.FaceParser (BatchNorm folded, auxiliary heads dropped) against the reference's own BiSeNet class
(oracle/_ref/Util/face_parsing/BiSeNet.py, staged byte-identical by oracle/stage_ref.py) on CPU in fp64."""
import importlib
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle', '_ref')


def _reference_bisenet():
    if not os.path.exists(os.path.join(REF, 'Util', 'face_parsing', 'BiSeNet.py')):
        pytest.skip('reference not staged (oracle/stage_ref.py)')
    import torch.utils.model_zoo as mz
    real = mz.load_url
    mz.load_url = lambda *a, **k: {}             # resnet.py:83 downloads ResNet18 weights: keep the random init
    sys.path.insert(0, REF)
    try:
        for k in [k for k in sys.modules if k == 'Util' or k.startswith('Util.')]:
            del sys.modules[k]
        mod = importlib.import_module('Util.face_parsing.BiSeNet')
        torch.manual_seed(3)
        net = mod.BiSeNet(n_classes=19)
    finally:
        mz.load_url = real
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'Util' or k.startswith('Util.')]:
            del sys.modules[k]
    return net


def test_face_parser_matches_reference_bisenet():
    from b200gan.parsing import FaceParser, synthetic_state_dict
    net = _reference_bisenet()
    # non-trivial BatchNorm statistics (a fresh BatchNorm is the identity and would hide a folding bug)
    sd = net.state_dict()
    syn = synthetic_state_dict(5)
    assert {k for k in sd if 'num_batches' not in k and not k.startswith(('conv_out16', 'conv_out32'))} == set(syn), \
        'synthetic_state_dict must mirror the reference key set (minus the auxiliary heads)'
    for k, v in syn.items():
        assert tuple(v.shape) == tuple(sd[k].shape), k
        sd[k].copy_(v)
    net = net.double().eval()
    fp = FaceParser.from_state_dict(net.state_dict()).double()
    x = torch.randn(2, 3, 96, 64, dtype=torch.float64)
    with torch.no_grad():
        ref = net(x)[0]
        got = fp(x)
    assert got[1] is None and got[2] is None
    err = float((got[0] - ref).abs().max() / ref.abs().max())
    assert err <= 2e-6, err            # the folded weights are stored in fp32
    assert float((got[0].argmax(1) != ref.argmax(1)).double().mean()) <= 1e-4
