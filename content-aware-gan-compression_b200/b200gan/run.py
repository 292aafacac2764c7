"""Compat launcher: run an UNMODIFIED reference driver script against the drop-in `model` / `op` modules.

    python -m b200gan.run [launcher options] /path/to/reference/prune.py [script options]

"Unmodified" means the script text is executed as is (`runpy.run_path(..., run_name='__main__')`).  The launcher
only prepares the process around it:

  1. `model` and `op` are imported from THIS package and pinned in `sys.modules` before the script (and its
     `Util/`, `Evaluation/` helpers, which are taken from the script's own directory) import them by name
     (train.py:15, Util/network_util.py:8, Evaluation/ppl.py:10);
  2. environment drift of this image is shimmed -- none of it concerns our kernels:
       * torchvision.utils.make_grid / save_image lost the `range=` keyword (Util/network_util.py:46-47, train.py:428-434);
       * numpy >= 1.24 refuses ragged `np.array([...])` (prune.py:45); scipy >= 1.16 dropped `sqrtm(disp=)`
         (Evaluation/fid.py:42);
       * `import lpips` needs `skimage` / `IPython` (lpips/__init__.py:7, lpips/networks_basic.py:11-12);
       * pretrained weights are downloaded (Util/face_parsing/resnet.py:83, lpips/pretrained_networks.py:100,
         Evaluation/inception.py:188) and there is no network: downloads return an empty marker and the module keeps its
         random initialisation (BiSeNet's full weights come from the vendored 79999_iter.pth afterwards);
       * hard-coded `gpu_device_ids = [0, 1]` (get_fid.py:10, get_ppl.py:10, train_hyperparams.py:14): nn.DataParallel
         device lists are clamped to the devices that exist;
       * the 66 MB FFHQ Inception statistics may be absent from a staged copy: `--synthetic-fid-stats` supplies
         zero-mean / identity-covariance statistics (the FID VALUE is then meaningless; the sampling loop is the path);
  3. optional, for parity tests: `--seed S` seeds torch / numpy / random; `--synthetic-mask` replaces the BiSeNet
     face parser by the centred ellipse used by bench.py and the tests (a random-init generator draws no faces).

  4. the rest of the KD loss (train.py only, `--kd-loss own`, the default): `lpips.PerceptualLoss` objects route CUDA
     calls of the net-lin VGG16 distance to b200gan.lpips.PerceptualLossVGG (parameters taken from the reference object),
     and `Batch_Img_Parsing` / `Get_Masked_Tensor` (Util/content_aware_pruning.py:61-117) are replaced by the device-side
     glue of b200gan.maskglue (same results, no host round trip); `--kd-loss reference` leaves both untouched.

The convolution engine is the package default (`CAGC_CONV_ALGO`, tcgen05 TF32 on sm_100); the saliency pass
(`Get_Content_Aware_Pruning_Score`) is wrapped in `config.exact_fp32()` unless `--saliency-engine` says otherwise, because
its product -- the prune mask -- is gated bit-exact (SURVEY.md finding 7).
"""
from __future__ import annotations

import argparse
import os
import pickle
import random
import runpy
import sys
import types


class OfflineWeights(dict):
    """Marker returned instead of a downloaded state_dict: load_state_dict(OfflineWeights) keeps the random init."""


def _shim_torchvision_range():
    import torchvision.utils as vu

    def wrap(fn):
        def inner(*a, **k):
            if 'range' in k:
                k['value_range'] = k.pop('range')
            return fn(*a, **k)
        inner._cagc_range_shim = True
        return inner
    for name in ('make_grid', 'save_image'):
        fn = getattr(vu, name)
        if not getattr(fn, '_cagc_range_shim', False):
            setattr(vu, name, wrap(fn))


def _shim_scipy_sqrtm():
    """scipy >= 1.16 dropped sqrtm(disp=...) (Evaluation/fid.py:42 passes disp=False and unpacks (sqrt, errest))."""
    try:
        from scipy import linalg
    except Exception:
        return
    real = linalg.sqrtm
    if getattr(real, '_cagc_disp_shim', False):
        return
    import inspect
    try:
        if 'disp' in inspect.signature(real).parameters:
            return
    except (TypeError, ValueError):
        pass

    def sqrtm(A, disp=True, **k):
        out = real(A, **k)
        return out if disp else (out, 0.0)
    sqrtm._cagc_disp_shim = True
    linalg.sqrtm = sqrtm


def _shim_ragged_numpy():
    import numpy as np
    real = np.array
    if getattr(real, '_cagc_ragged', False):
        return

    def array(obj, *a, **k):
        try:
            return real(obj, *a, **k)
        except ValueError as e:
            if 'inhomogeneous' in str(e) and 'dtype' not in k and not a:
                return real(obj, dtype=object, **k)       # what numpy < 1.24 did implicitly (prune.py:45)
            raise
    array._cagc_ragged = True
    np.array = array


def _shim_missing_modules():
    """`import lpips` pulls in skimage (measure.compare_ssim, color, transform) and IPython.embed at import time
    (lpips/__init__.py:7, lpips/dist_model.py:16, lpips/networks_basic.py:11-12); none of them is used by the
    'net-lin' VGG distance that train.py / Evaluation/ppl.py call.  Missing ones become empty stub modules."""
    def stub(name, is_pkg=False, **attrs):
        if name in sys.modules:
            return
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            if is_pkg:
                m.__path__ = []
            m.__dict__.update(attrs)
            sys.modules[name] = m
            if '.' in name:
                parent, child = name.rsplit('.', 1)
                setattr(sys.modules[parent], child, m)

    def unavailable(*a, **k):
        raise RuntimeError('scikit-image is not installed in this image')
    stub('skimage', is_pkg=True)
    stub('skimage.measure', compare_ssim=unavailable)
    stub('skimage.color')
    stub('skimage.transform')
    if not hasattr(sys.modules['skimage.measure'], 'compare_ssim'):      # real, newer scikit-image renamed it
        sys.modules['skimage.measure'].compare_ssim = getattr(sys.modules['skimage.measure'], 'structural_similarity',
                                                              unavailable)
    stub('IPython', embed=lambda *a, **k: None)


def _shim_downloads():
    import torch
    import torch.hub
    import torch.utils.model_zoo as mz

    def offline(url, *a, **k):
        print(f'[b200gan.run] offline: not downloading {url}; module keeps its random initialisation', file=sys.stderr)
        return OfflineWeights()
    torch.hub.load_state_dict_from_url = offline
    mz.load_url = offline
    for modname in ('torchvision._internally_replaced_utils', 'torchvision.models._api', 'torchvision.models.utils'):
        try:
            m = __import__(modname, fromlist=['x'])
            if hasattr(m, 'load_state_dict_from_url'):
                m.load_state_dict_from_url = offline
        except Exception:
            pass
    real = torch.nn.Module.load_state_dict
    if not getattr(real, '_cagc_offline', False):
        def load_state_dict(self, state_dict, *a, **k):
            if isinstance(state_dict, OfflineWeights):
                return torch.nn.modules.module._IncompatibleKeys([], [])
            return real(self, state_dict, *a, **k)
        load_state_dict._cagc_offline = True
        torch.nn.Module.load_state_dict = load_state_dict


def _shim_dataparallel_devices():
    import torch
    real = torch.nn.DataParallel.__init__
    if getattr(real, '_cagc_clamped', False):
        return

    def init(self, module, device_ids=None, output_device=None, dim=0):
        n = torch.cuda.device_count()
        if device_ids is not None:
            ids = [d for d in device_ids if (d.index if isinstance(d, torch.device) else int(d)) < n]
            if ids != list(device_ids):
                print(f'[b200gan.run] DataParallel device_ids {list(device_ids)} -> {ids} ({n} device(s) present)',
                      file=sys.stderr)
            device_ids = ids
        return real(self, module, device_ids=device_ids, output_device=output_device, dim=dim)
    init._cagc_clamped = True
    torch.nn.DataParallel.__init__ = init


def _synthetic_fid_stats(script_dir):
    import numpy as np
    d = os.path.join(script_dir, 'Evaluation', 'inception_ffhq_embed')
    for name in ('self_ffhq_256_inception_embeddings_eval_mode.pkl', 'self_ffhq_1024_inception_embeddings_eval_mode.pkl'):
        path = os.path.join(d, name)
        if not os.path.exists(path):
            os.makedirs(d, exist_ok=True)
            with open(path, 'wb') as f:
                pickle.dump({'mean': np.zeros(2048), 'cov': np.eye(2048)}, f)
            print(f'[b200gan.run] wrote synthetic Inception statistics {path}', file=sys.stderr)


def _ellipse(size):
    import numpy as np
    yy, xx = np.mgrid[0:size, 0:size]
    c = (size - 1) / 2
    return (((yy - c) / (0.42 * size)) ** 2 + ((xx - c) / (0.34 * size)) ** 2) <= 1


def _patch_saliency(synthetic_mask: bool, engine: str):
    """Hooks on the reference's Util.content_aware_pruning (imported from the script's directory)."""
    import numpy as np
    from b200gan import config
    import Util.content_aware_pruning as cap
    if synthetic_mask:
        # A stand-in for BiSeNet with the same call contract (`parsing_net(x)[0]` = class scores [B,19,H,W]): class 1
        # inside the centred ellipse, class 0 outside.  Extract_Face_Mask (:38-58), Batch_Img_Parsing (:61-88) and
        # Get_Masked_Tensor (:90-117) then run unmodified on it.
        import torch
        from torchvision import transforms

        class SyntheticParser:
            def __call__(self, x):
                b, _, h, w = x.shape
                inside = torch.from_numpy(_ellipse(h)).to(x.device)
                out = torch.zeros(b, 19, h, w, device=x.device)
                out[:, 1] = inside
                out[:, 0] = ~inside
                return (out,)

            def eval(self):
                return self

            def to(self, *a, **k):
                return self

        to_tensor = transforms.Compose([transforms.ToTensor(),
                                        transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))])
        cap.Get_Parsing_Net = lambda device: (SyntheticParser(), to_tensor)
    if engine != 'default':
        algo = {'fp32': config.ALGO_SIMT_FP32, 'tf32': config.ALGO_TCGEN05_TF32,
                '3xtf32': getattr(config, 'ALGO_TCGEN05_3XTF32', config.ALGO_SIMT_FP32)}[engine]
        real = cap.Get_Content_Aware_Pruning_Score

        def scored(*a, **k):
            with config.use_algo(algo):
                return real(*a, **k)
        cap.Get_Content_Aware_Pruning_Score = scored


def _patch_kd_loss():
    """train.py binds `Batch_Img_Parsing` / `Get_Masked_Tensor` by name at import (train.py:22) and builds
    `lpips.PerceptualLoss(model='net-lin', net='vgg', ...)` (train.py:510): both are swapped at their source modules."""
    import torch
    import Util.content_aware_pruning as cap
    from b200gan import maskglue
    from b200gan.lpips import PerceptualLossVGG
    ref_parsing, ref_masked = cap.Batch_Img_Parsing, cap.Get_Masked_Tensor

    def Batch_Img_Parsing(img_tensor, parsing_net, device):
        if not img_tensor.is_cuda:
            return ref_parsing(img_tensor, parsing_net, device)
        with torch.no_grad():
            return parsing_net(maskglue.parse_preprocess(img_tensor))[0].argmax(1)       # labels [N,512,512], as :87

    def Get_Masked_Tensor(img_tensor, batch_parsing, device, mask_grad=False):
        if not img_tensor.is_cuda:
            return ref_masked(img_tensor, batch_parsing, device, mask_grad)
        labels = batch_parsing.to(img_tensor.device).float().unsqueeze(1)               # K = 1: the plane holds labels
        return img_tensor * maskglue.parsing_mask(labels, img_tensor.shape[-1])
    cap.Batch_Img_Parsing, cap.Get_Masked_Tensor = Batch_Img_Parsing, Get_Masked_Tensor

    import lpips
    Real = lpips.PerceptualLoss
    if getattr(Real, '_cagc_swapped', False):
        return
    # methods are patched in place: PerceptualLoss.__init__ calls `super(PerceptualLoss, self)` through the module-global
    # name (lpips/__init__.py:17), so rebinding that name to a subclass would recurse
    real_init, real_forward = Real.__init__, Real.forward

    def __init__(self, *a, **k):
        real_init(self, *a, **k)
        try:
            own = PerceptualLossVGG.from_reference(self)
        except (ValueError, AttributeError):
            own = None                          # alex / squeeze / spatial / L2 variants stay on the reference code
        object.__setattr__(self, '_cagc_own', own)

    def forward(self, pred, target, normalize=False):
        own = getattr(self, '_cagc_own', None)
        if own is not None and pred.is_cuda and pred.dtype == torch.float32 and pred.shape[-1] % 16 == 0 \
                and pred.shape[-2] % 16 == 0:
            if own.conv_weights[0].device != pred.device:
                own.to(pred.device)
            return own(pred, target, normalize)
        return real_forward(self, pred, target, normalize)
    Real.__init__, Real.forward, Real._cagc_swapped = __init__, forward, True


def main(argv=None):
    ap = argparse.ArgumentParser(prog='python -m b200gan.run', description=__doc__.split('\n\n')[0])
    ap.add_argument('--seed', type=int, default=None)
    ap.add_argument('--synthetic-mask', action='store_true')
    ap.add_argument('--synthetic-fid-stats', action='store_true')
    ap.add_argument('--saliency-engine', default='fp32', choices=['fp32', 'tf32', '3xtf32', 'default'])
    ap.add_argument('--kd-loss', default='own', choices=['own', 'reference'],
                    help="train.py: LPIPS / content-mask glue on this package's kernels (own) or the reference's code")
    ap.add_argument('--cwd', default=None, help='working directory for the script (its relative ./Model paths)')
    ap.add_argument('script')
    ap.add_argument('script_args', nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)

    script = os.path.abspath(args.script)
    script_dir = os.path.dirname(script)
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # our `model` / `op` first, then the script's own tree (Util/, Evaluation/, lpips/, train_hyperparams.py, ...)
    sys.path[:] = [pkg, script_dir] + [p for p in sys.path if p not in (pkg, script_dir, '')]
    import torch
    import model
    import op
    assert os.path.dirname(os.path.abspath(model.__file__)) == pkg and op.__file__.startswith(pkg), \
        'the drop-in model/op must be the ones bound by name'
    _shim_torchvision_range()
    _shim_ragged_numpy()
    _shim_scipy_sqrtm()
    _shim_missing_modules()
    _shim_downloads()
    _shim_dataparallel_devices()
    if args.synthetic_fid_stats:
        _synthetic_fid_stats(script_dir)
    name = os.path.basename(script)
    if name in ('prune.py', 'train.py') or args.synthetic_mask:
        _patch_saliency(args.synthetic_mask, args.saliency_engine)
    if name == 'train.py' and args.kd_loss == 'own':
        _patch_kd_loss()
    if args.seed is not None:
        import numpy as np
        torch.manual_seed(args.seed)
        np.random.seed(args.seed)
        random.seed(args.seed)
    if args.cwd:
        os.makedirs(args.cwd, exist_ok=True)
        os.chdir(args.cwd)
    sys.argv = [script] + list(args.script_args)
    runpy.run_path(script, run_name='__main__')


if __name__ == '__main__':
    main()
