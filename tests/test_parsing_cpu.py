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


def test_perceptual_loss_vgg_from_reference_object():
    """PerceptualLossVGG.from_reference takes the 13 VGG16 convolutions, the five lin heads and the scaling-layer
    constants from the object the reference's train.py:510 builds (oracle/_ref/lpips, random VGG16: no download)."""
    import types
    if not os.path.exists(os.path.join(REF, 'lpips', '__init__.py')):
        pytest.skip('reference not staged (oracle/stage_ref.py)')
    import torchvision.models as tvm
    real_vgg = tvm.vgg16
    stubs = []
    for name in ('skimage', 'skimage.measure', 'skimage.color', 'skimage.transform', 'IPython'):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__path__ = []
                sys.modules[name] = m
                stubs.append(name)
    sys.modules['skimage.measure'].__dict__.setdefault('compare_ssim', None)
    sys.modules['IPython'].__dict__.setdefault('embed', lambda *a, **k: None)
    tvm.vgg16 = lambda pretrained=True, **k: real_vgg(weights=None)
    sys.path.insert(0, REF)
    try:
        import lpips as ref_lpips
        assert os.path.realpath(ref_lpips.__file__).startswith(os.path.realpath(REF))
        torch.manual_seed(11)
        ref = ref_lpips.PerceptualLoss(model='net-lin', net='vgg', use_gpu=False)
    finally:
        tvm.vgg16 = real_vgg
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'lpips' or k.startswith('lpips.')] + stubs:
            sys.modules.pop(k, None)
    from b200gan.lpips import PerceptualLossVGG, VGG16_CFG
    own = PerceptualLossVGG.from_reference(ref)
    net = ref.model.net
    convs = [m for sl in (net.net.slice1, net.net.slice2, net.net.slice3, net.net.slice4, net.net.slice5) for m in sl
             if isinstance(m, torch.nn.Conv2d)]
    assert len(convs) == 13 == len(own.conv_weights)
    assert [w.shape[0] for w in own.conv_weights] == [c for c in VGG16_CFG if c != 'M']
    for m, w, b in zip(convs, own.conv_weights, own.conv_biases):
        assert torch.equal(m.weight, w) and torch.equal(m.bias, b) and not w.requires_grad
    for lin, w in zip((net.lin0, net.lin1, net.lin2, net.lin3, net.lin4), own.lin_weights):
        assert torch.equal(lin.model[1].weight.reshape(-1), w) and float(w.min()) >= 0      # the vendored heads are >= 0
    assert [round(v, 3) for v in own._shift] == [-0.03, -0.088, -0.188]
    assert [round(v, 3) for v in own._scale] == [0.458, 0.448, 0.45]
    # product code is CUDA-only: a CPU call fails loudly instead of falling back
    with pytest.raises(RuntimeError, match='CUDA-only'):
        own(torch.zeros(1, 3, 32, 32), torch.zeros(1, 3, 32, 32))
