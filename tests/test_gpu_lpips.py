"""Rest of the KD loss on the package's own kernels (b200gan/lpips.py, b200gan/maskglue.py, csrc/lpips.cu, csrc/kdloss.cu)
against the fp64 CPU oracle (oracle/lpips_oracle.py, pinned to the reference's own lpips / Util code by
tests/test_oracle_vs_golden.py) and against the reference-generated fixtures directly."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import synth

pytestmark = pytest.mark.gpu


def _t(x):
    return torch.from_numpy(np.asarray(x)) if not torch.is_tensor(x) else x


def relmax(a, b):
    a, b = _t(a).detach().double().cpu(), _t(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a, b):
    a, b = _t(a).detach().double().cpu(), _t(b).detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def median_rel(a, b):
    """Median element-wise relative error: insensitive to the few elements a flipped ReLU mask moves."""
    a, b = _t(a).detach().double().cpu().flatten(), _t(b).detach().double().cpu().flatten()
    keep = b.abs() > 1e-3 * b.abs().max()
    return float(((a - b).abs()[keep] / b.abs()[keep]).median())


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous().float().cuda()


def _nchw(buf):
    return buf.permute(0, 3, 1, 2)


def _st():
    return torch.cuda.current_stream().cuda_stream


def _f3(v):
    return (C.c_float * 3)(*v)


@pytest.mark.parametrize('h,w,cout', [(16, 16, 64), (9, 21, 64), (40, 24, 32)])
def test_rgb_conv3x3_fwd_bwd(h, w, cout):
    """conv1_1 with the scaling layer folded in, against F.conv2d in fp64 (image read through strides)."""
    from b200gan._lib import lib, check
    g = torch.Generator().manual_seed(h * 31 + w)
    img = torch.randn(3, 3, h, w, generator=g, dtype=torch.float64)
    wt = torch.randn(cout, 3, 3, 3, generator=g, dtype=torch.float64) * 0.3
    bias = torch.randn(cout, generator=g, dtype=torch.float64) * 0.2
    shift, scale = (-.030, -.088, -.188), (.458, .448, .450)
    img_leaf = img.clone().requires_grad_(True)
    pre = F.conv2d((img_leaf - torch.tensor(shift).view(1, 3, 1, 1)) / torch.tensor(scale).view(1, 3, 1, 1), wt, bias, padding=1)
    ref = F.relu(pre)
    # a strided view: channels-last image
    img_cl = img.float().cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    out = torch.empty((3, h, w, cout), device='cuda')
    sb, sc, sh, sw = img_cl.stride()
    wd, bd = wt.float().cuda(), bias.float().cuda()      # named: a temporary's block would be recycled before the launch runs
    check(lib.cagc_rgb_conv3x3_fwd(_st(), img_cl.data_ptr(), sb, sc, sh, sw, wd.data_ptr(), bd.data_ptr(), _f3(shift),
                                   _f3(scale), out.data_ptr(), 3, h, w, cout, cout))
    assert relmax(_nchw(out), ref) <= 5e-6
    gz = torch.randn(3, cout, h, w, generator=g, dtype=torch.float64)
    gref, = torch.autograd.grad(pre, img_leaf, gz)
    gimg = torch.empty((3, 3, h, w), device='cuda')
    gzd = _nhwc(gz)
    check(lib.cagc_rgb_conv3x3_bwd(_st(), gzd.data_ptr(), wd.data_ptr(), _f3(scale), gimg.data_ptr(), 3, h, w, cout, cout))
    assert relmax(gimg, gref) <= 5e-6


@pytest.mark.parametrize('h,w,c', [(8, 8, 64), (6, 10, 40), (32, 32, 128)])
def test_maxpool_and_relu_pool_bwd(h, w, c):
    """MaxPool2d(2,2) forward; backward of `max_pool2d(a)` summed with a direct gradient and masked by a > 0, where a is
    a post-ReLU activation with exact zeros (ties) -- against autograd of relu -> max_pool2d in fp64."""
    from b200gan._lib import lib, check
    g = torch.Generator().manual_seed(h * 7 + c)
    z = torch.randn(2, c, h, w, generator=g, dtype=torch.float64).float().double().requires_grad_(True)
    a = F.relu(z)
    pooled = F.max_pool2d(a, 2, 2)
    ab = _nhwc(a)
    out = torch.empty((2, h // 2, w // 2, c), device='cuda')
    check(lib.cagc_maxpool2_nhwc(_st(), ab.data_ptr(), out.data_ptr(), 2, h, w, c))
    assert torch.equal(_nchw(out).cpu().double(), pooled.detach())
    gp = torch.randn(pooled.shape, generator=g, dtype=torch.float64)
    gd = torch.randn(a.shape, generator=g, dtype=torch.float64)
    gref, = torch.autograd.grad([pooled, a], z, [gp, gd], retain_graph=True)
    gz = torch.empty((2, h, w, c), device='cuda')
    gpd, gdd = _nhwc(gp), _nhwc(gd)
    check(lib.cagc_relu_pool_bwd(_st(), ab.data_ptr(), gpd.data_ptr(), gdd.data_ptr(), gz.data_ptr(), 2, h, w, c))
    assert relmax(_nchw(gz), gref) <= 1e-6
    # pool gradient only / direct gradient only
    gref_p, = torch.autograd.grad(pooled, z, gp, retain_graph=True)
    check(lib.cagc_relu_pool_bwd(_st(), ab.data_ptr(), gpd.data_ptr(), None, gz.data_ptr(), 2, h, w, c))
    assert relmax(_nchw(gz), gref_p) <= 1e-6
    gref_d, = torch.autograd.grad(a, z, gd)
    check(lib.cagc_relu_pool_bwd(_st(), ab.data_ptr(), None, gdd.data_ptr(), gz.data_ptr(), 2, h, w, c))
    assert relmax(_nchw(gz), gref_d) <= 1e-6


@pytest.mark.parametrize('c,hw', [(64, 100), (128, 64), (256, 33), (512, 16), (512, 1)])
def test_lpips_head_fwd_bwd(c, hw):
    """One tap of networks_basic.py:66-84 against the oracle formulation in fp64 (features are post-ReLU: >= 0)."""
    from b200gan._lib import lib, check
    from oracle import lpips_oracle as L
    g = torch.Generator().manual_seed(c + hw)
    n = 3
    fs = F.relu(torch.randn(n, c, hw, 1, generator=g, dtype=torch.float64)).requires_grad_(True)
    ft = F.relu(torch.randn(n, c, hw, 1, generator=g, dtype=torch.float64))
    lw = torch.rand(c, generator=g, dtype=torch.float64)
    d = (L.normalize_tensor(ft) - L.normalize_tensor(fs)) ** 2
    val = F.conv2d(d, lw.view(1, c, 1, 1)).mean([2, 3]).reshape(n)
    gv = torch.randn(n, generator=g, dtype=torch.float64)
    gref, = torch.autograd.grad(val, fs, gv)
    fsb, ftb = _nhwc(fs), _nhwc(ft)
    nblk = int(lib.cagc_lpips_head_blocks(n, hw))
    partial = torch.empty(n * nblk, device='cuda')
    out = torch.full((n,), 7.0, device='cuda')
    lwc = lw.float().cuda()
    check(lib.cagc_lpips_head_fwd(_st(), fsb.data_ptr(), ftb.data_ptr(), lwc.data_ptr(), partial.data_ptr(), out.data_ptr(),
                                  n, hw, c, 0))
    assert relmax(out, val) <= 1e-5
    check(lib.cagc_lpips_head_fwd(_st(), fsb.data_ptr(), ftb.data_ptr(), lwc.data_ptr(), partial.data_ptr(), out.data_ptr(),
                                  n, hw, c, 1))
    assert relmax(out, 2 * val) <= 1e-5
    gs = torch.empty_like(fsb)
    gvd = gv.float().cuda()
    check(lib.cagc_lpips_head_bwd(_st(), fsb.data_ptr(), ftb.data_ptr(), lwc.data_ptr(), gvd.data_ptr(), gs.data_ptr(), n, hw, c))
    assert relmax(_nchw(gs), gref) <= 2e-5


def _lpips_module(seed, lins):
    from b200gan.lpips import PerceptualLossVGG
    cw, cb = synth.vgg16_weights(seed)
    return PerceptualLossVGG([torch.from_numpy(a) for a in cw], [torch.from_numpy(a) for a in cb],
                             [torch.from_numpy(np.asarray(a)) for a in lins]).cuda(), cw, cb


# Gradient tolerances are L2: the gradient passes 13 ReLU masks, and a pre-activation within rounding distance of zero flips
# its mask, which moves single elements by their full magnitude -- a fraction f of flipped units costs sqrt(f) in relative
# L2 (fp32: f ~ 1e-6 -> 1e-3 once a layer has millions of units; TF32: 13 layers of 2^-10 operand rounding on top).
@pytest.mark.parametrize('algo,tol,gtol', [(0, 2e-5, 2e-4), (1, 5e-3, 8e-2)])
def test_lpips_matches_reference_golden(golden_dir, algo, tol, gtol):
    """PerceptualLossVGG on both engines against the distances and the pred-gradient the REFERENCE's own
    lpips.PerceptualLoss produced in fp64 (tests/golden/lpips_tiny.npz), 32x32 images."""
    from b200gan import config
    g = np.load(os.path.join(golden_dir, 'lpips_tiny.npz'))
    c = synth.LPIPS_TINY
    mod, _, _ = _lpips_module(c['seed_weights'], [g[f'lin{k}'] for k in range(5)])
    pred_np, target_np = synth.lpips_images(c['seed_inputs'], c['batch'], c['size'])
    pred = torch.from_numpy(pred_np).float().cuda().requires_grad_(True)
    target = torch.from_numpy(target_np).float().cuda()
    with config.use_algo(algo):
        val = mod(pred, target)
        assert tuple(val.shape) == (c['batch'], 1, 1, 1)
        gp, = torch.autograd.grad(val, pred, torch.from_numpy(g['cot']).float().cuda())
    assert relmax(val, torch.from_numpy(g['val'])) <= tol
    assert rel_l2(gp, torch.from_numpy(g['g_pred'])) <= gtol
    with config.use_algo(algo):
        kd = 3.0 * torch.mean(mod(pred, target))              # train.py:182
        g2, = torch.autograd.grad(kd, pred)
    assert abs(float(kd) - float(g['kd_lpips'])) <= tol * abs(float(g['kd_lpips']))
    assert rel_l2(g2, torch.from_numpy(g['kd_lpips_g'])) <= gtol


@pytest.mark.parametrize('algo,tol,gtol,mtol', [(0, 5e-5, 1e-2, 2e-5), (1, 5e-3, 8e-2, 5e-2)])
@pytest.mark.parametrize('size,batch', [(64, 3), (256, 2)])
def test_lpips_matches_oracle(golden_dir, algo, tol, gtol, mtol, size, batch):
    """Larger images (256px = BASELINE configs[1]) against the fp64 oracle; pred arrives as a non-contiguous view."""
    from b200gan import config
    from oracle import lpips_oracle as L
    g = np.load(os.path.join(golden_dir, 'lpips_tiny.npz'))
    lins = [g[f'lin{k}'] for k in range(5)]
    mod, cw, cb = _lpips_module(61, lins)
    pred_np, target_np = synth.lpips_images(62 + size, batch, size)
    pred64 = torch.from_numpy(pred_np).requires_grad_(True)
    ref = L.lpips_vgg(pred64, torch.from_numpy(target_np), [torch.from_numpy(a) for a in cw],
                      [torch.from_numpy(a) for a in cb], [torch.from_numpy(np.asarray(a)).double() for a in lins])
    gref, = torch.autograd.grad(ref.sum(), pred64)
    pred = torch.from_numpy(pred_np).float().cuda().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2).requires_grad_(True)
    target = torch.from_numpy(target_np).float().cuda()
    with config.use_algo(algo):
        val = mod(pred, target)
        gp, = torch.autograd.grad(val.sum(), pred)
    assert relmax(val, ref) <= tol
    assert rel_l2(gp, gref) <= gtol           # dominated by single flipped ReLU masks (see above) ...
    assert median_rel(gp, gref) <= mtol       # ... the typical element agrees to rounding
    # target does not receive a gradient (the reference flags the teacher image requires_grad, train.py:160, but never
    # reads that gradient)
    t2 = target.clone().requires_grad_(True)
    with config.use_algo(algo):
        mod(pred, t2).sum().backward()
    assert t2.grad is None


@pytest.mark.parametrize('tag,size', [('s256', 256), ('s1024', 1024), ('s64', 64)])
def test_mask_glue_matches_reference(golden_dir, tag, size):
    """parse_preprocess / parsing_mask against what the reference's Batch_Img_Parsing fed its parser and the mask its
    Get_Masked_Tensor applied (tests/golden/mask_glue.npz): preprocessing <= 1e-5, mask bit-exact."""
    from b200gan import maskglue
    g = np.load(os.path.join(golden_dir, 'mask_glue.npz'))
    n = 2
    rs = np.random.RandomState(910 + size)
    img = torch.from_numpy((rs.standard_normal((n, 3, size, size)) * 0.8).astype(np.float32)).cuda()
    pre = maskglue.parse_preprocess(img)
    assert tuple(pre.shape) == (n, 3, 512, 512)
    assert relmax(pre[:, :, ::7, ::5], torch.from_numpy(g[f'{tag}.pre'])) <= 1e-5
    # strided input (channels-last image)
    pre2 = maskglue.parse_preprocess(img.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2))
    assert torch.equal(pre, pre2)
    pre3 = maskglue.parse_preprocess(img, channels_last=True)
    assert pre3.is_contiguous(memory_format=torch.channels_last) and torch.equal(pre, pre3)
    scores = torch.from_numpy(synth.parser_scores(911 + size, n)).cuda()
    mask = maskglue.parsing_mask(scores, size)
    want = np.unpackbits(g[f'{tag}.mask'])[:n * size * size].reshape(n, 1, size, size)
    assert np.array_equal(mask.cpu().numpy().astype(np.uint8), want)

    class Parser:
        def __call__(self, x):
            assert torch.equal(x, pre)
            return (scores,)
    assert torch.equal(maskglue.content_mask(img, Parser()), mask)
    masked = img * mask
    assert abs(float(masked.double().abs().sum()) - float(g[f'{tag}.masked_sum'])) <= 1e-5 * float(g[f'{tag}.masked_sum'])


@pytest.mark.parametrize('mode', ['Output_Only', 'Intermediate'])
def test_kd_step_full_loss_vs_reference_golden(golden_dir, mode):
    """KDStep with the parsed content mask and LPIPS (the complete generator loss of train.py:280-308) against the fixture
    produced by the reference's own KD_loss text + lpips.PerceptualLoss + Batch_Img_Parsing / Get_Masked_Tensor: total
    loss and every student gradient, exact-fp32 engine."""
    from test_gpu_baseline_configs import _kd_tiny
    from b200gan import config
    from b200gan.kd import KDStep
    c, _, disc, student, teacher, z, s_noise, t_noise = _kd_tiny(golden_dir)
    g = np.load(os.path.join(golden_dir, 'kd_full_tiny.npz'))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    mod, _, _ = _lpips_module(synth.KD_FULL['seed_vgg'], [g[f'lin{k}'] for k in range(5)])
    scores = torch.from_numpy(synth.parser_scores(synth.KD_FULL['seed_parser'], c['batch'])).cuda()
    kd = KDStep(student, teacher, disc, kd_mode=mode, percept_loss=mod, parsing_net=lambda x: (scores,))
    with config.exact_fp32():
        loss = kd.step(z, c['inject'], s_noise, t_noise)
    ref_loss = float(g[f'{mode}.g_loss']) + float(g[f'{mode}.kd_l1']) + float(g[f'{mode}.kd_lpips'])
    assert abs(float(loss) - ref_loss) <= 1e-4 * abs(ref_loss), (float(loss), ref_loss)
    for n, p in student.named_parameters():
        ref = g[f'{mode}.grad.{n}']
        if np.abs(ref).max() == 0:
            assert float(p.grad.abs().max()) == 0.0, n
            continue
        e = relmax(p.grad, ref)
        assert e <= 1e-3, f'{mode} grad {n}: {e:.3e}'


def test_kd_step_with_lpips_captures_into_a_graph(golden_dir):
    """The complete step (parser glue and LPIPS included) has no host round trip: it captures into one CUDA graph and the
    replay reproduces the eager loss on the same latents."""
    import copy
    from test_gpu_baseline_configs import _kd_tiny
    from b200gan import config
    from b200gan.kd import KDStep
    c, _, disc, student, teacher, z, _, _ = _kd_tiny(golden_dir)
    g = np.load(os.path.join(golden_dir, 'kd_full_tiny.npz'))
    mod, _, _ = _lpips_module(synth.KD_FULL['seed_vgg'], [g[f'lin{k}'] for k in range(5)])
    scores = torch.from_numpy(synth.parser_scores(synth.KD_FULL['seed_parser'], c['batch'])).cuda()
    with torch.no_grad():                       # deterministic step: no per-layer noise
        for m in (student, teacher):
            for n, p in m.named_parameters():
                if n.endswith('noise.weight'):
                    p.zero_()
    student2 = copy.deepcopy(student)
    with config.use_algo(config.ALGO_TCGEN05_TF32):
        eager = KDStep(student, teacher, disc, percept_loss=mod, parsing_net=lambda x: (scores,), lr=0.0)
        l_eager = float(eager.step(z, c['inject']))
        graphed = KDStep(student2, teacher, disc, percept_loss=mod, parsing_net=lambda x: (scores,), lr=0.0)
        graphed.capture(c['batch'], c['inject'], style_dim=c['style_dim'])
        l_graph = float(graphed.step_graphed(z))
    assert abs(l_eager - l_graph) <= 1e-5 * abs(l_eager), (l_eager, l_graph)


def test_face_parser_gpu_forms_agree():
    """FaceParser on the GPU: the fused cuDNN calls (conv + bias [+ residual] + ReLU) against the plain conv -> add -> relu
    composition in fp32 without TF32, and the low-resolution mask path (upsample inside the mask kernel) against
    materialising the 512^2 scores."""
    from b200gan import maskglue
    from b200gan.parsing import FaceParser, synthetic_state_dict
    torch.backends.cudnn.allow_tf32 = False
    fp = FaceParser.from_state_dict(synthetic_state_dict(7)).cuda()
    img = torch.tanh(torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(1))).cuda()
    x = maskglue.parse_preprocess(img)
    with torch.no_grad():
        fp.fused = True
        a = fp(x)[0]
        lr = fp.scores_lowres(x)
        fp.fused = False
        b = fp(x)[0]
    assert tuple(a.shape) == (2, 19, 512, 512) and tuple(lr.shape) == (2, 19, 64, 64)
    assert relmax(a, b) <= 1e-4
    m_full = maskglue.parsing_mask(a, 256)
    fp.fused = True
    m_lr = maskglue.content_mask(img, fp)
    assert tuple(m_lr.shape) == (2, 1, 256, 256)
    # same arithmetic up to the order of fp32 operations in the upsample: near-ties of two classes may flip single pixels
    assert float((m_full != m_lr).float().mean()) <= 1e-3
    assert 0.02 < float(m_lr.mean()) < 0.999
    torch.backends.cudnn.allow_tf32 = True
