"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the Python surface mirrors the reference's module tree / state_dict contract, and
the product path refuses CPU tensors instead of falling back."""
import os
import re
import ctypes

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from b200gan import _lib
    hdr = open(os.path.join(ROOT, 'include', 'cagc_b200.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    declared = set(re.findall(r'\b(cagc_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations parsed'
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f'{name} declared in include/cagc_b200.h but not exported'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert raw.cagc_abi_version() == _lib.ABI_VERSION
    m = re.search(r'#define\s+CAGC_ABI_VERSION\s+(\d+)', hdr)
    assert int(m.group(1)) == _lib.ABI_VERSION


def test_pure_host_entry_points():
    from b200gan import _lib
    L = _lib.lib
    assert L.cagc_bias_grad_chunks(16) == 0 and L.cagc_bias_grad_chunks(65536) == 16
    assert L.cagc_act_bwd_chunks(4, 4) == 1 and L.cagc_act_bwd_chunks(256, 256) == 64
    s = L.cagc_conv_wgrad_splits(16, 256, 256, 40, 40, 3, 0)
    assert 1 <= s <= 256
    assert L.cagc_conv_wgrad_splits(16, 4, 4, 512, 512, 3, 0) <= 2


def test_state_dict_contract_matches_reference(golden_dir):
    """Key names AND order (Util/mask_util.py consumes keys positionally) == the reference's."""
    import model
    g = np.load(os.path.join(golden_dir, 'generator_tiny.npz'), allow_pickle=True)
    shape = [int(v) for v in g['net_shape']]
    gen = model.Generator(32, 32, 2, generator_net_shape=shape)
    assert list(gen.state_dict().keys()) == [str(k) for k in g['sd_keys']]
    for k, v in gen.state_dict().items():
        assert tuple(v.shape) == g[f'sd.{k}'].shape, k
    assert gen.n_latent == 8 and gen.num_layers == 7 and gen.size == 32
    assert gen.conv1.conv.weight.shape == (1, shape[1], shape[0], 3, 3)
    assert gen.conv1.conv.demodulate and not gen.to_rgb1.conv.demodulate
    assert callable(gen.conv1.conv.modulation) and gen.conv1.conv.out_channel == shape[1]
    assert hasattr(gen.input.input, 'device') and hasattr(gen.noises, 'noise_0')


def test_default_widths_and_discriminator_tree():
    import model
    gen = model.Generator(64, 512, 8)
    assert [c.conv.in_channel for c in [gen.conv1] + list(gen.convs)] == [512] * 9
    assert sum(p.numel() for p in gen.parameters()) > 2e7
    d = model.Discriminator(64)
    keys = list(d.state_dict().keys())
    assert keys[0] == 'convs.0.0.weight' and keys[1] == 'convs.0.1.bias'
    assert 'convs.1.conv2.0.kernel' in keys and 'convs.1.skip.1.weight' in keys
    assert keys[-1] == 'final_linear.1.bias'


def test_pad_arithmetic():
    import model
    up = model.Upsample([1, 3, 3, 1])
    assert up.pad == (2, 1) and float(up.kernel.sum()) == pytest.approx(4.0)
    assert model.Downsample([1, 3, 3, 1]).pad == (1, 1)
    assert model.ModulatedConv2d(4, 4, 3, 8, upsample=True).blur.pad == (1, 1)
    assert model.ConvLayer(4, 4, 3, downsample=True)[0].pad == (2, 2)
    assert model.ConvLayer(4, 4, 1, downsample=True, activate=False, bias=False)[0].pad == (1, 1)


def test_cpu_tensors_are_rejected_not_served():
    import model
    import op
    with pytest.raises(RuntimeError, match='CUDA-only'):
        op.fused_leaky_relu(torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match='CUDA-only'):
        op.upfirdn2d(torch.zeros(1, 1, 4, 4), torch.ones(2, 2))
    gen = model.Generator(8, 8, 1, generator_net_shape=[4, 4, 4, 4])
    with pytest.raises(RuntimeError, match='CUDA-only'):
        gen([torch.zeros(1, 8)])


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'content-aware-gan-compression_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f


def test_tensor_cache_is_keyed_by_identity_and_version():
    """Derived values of long-lived tensors (FIR taps, frozen-weight slabs) are cached per tensor OBJECT and version
    counter -- a data pointer is not a safe key (the allocator recycles addresses of freed temporaries)."""
    import gc
    import torch
    from b200gan._lib import TensorCache
    c = TensorCache()
    calls = []

    def make(tag):
        calls.append(tag)
        return len(calls)

    a = torch.zeros(4)
    assert c.get(a, 'k', lambda: make('a')) == 1
    assert c.get(a, 'k', lambda: make('a')) == 1 and calls == ['a']            # hit
    assert c.get(a, 'other', lambda: make('a2')) == 2                          # different tag
    a.add_(1)                                                                  # in-place update bumps the version
    assert c.get(a, 'k', lambda: make('a3')) == 3
    b = a.clone()                                                              # equal values, different object
    assert c.get(b, 'k', lambda: make('b')) == 4
    n = len(c._d)
    del b
    gc.collect()
    assert len(c._d) == n - 1                                                  # entries die with their tensor
    assert c.get(a, 'none', lambda: None) is None and ('id', 'none') not in c._d
    before = len(calls)
    a.data = torch.ones(4)                                                     # what module.to(device) does: same object,
    assert c.get(a, 'k', lambda: make('a4')) == before + 1                     # same version, new storage -> rebuilt
