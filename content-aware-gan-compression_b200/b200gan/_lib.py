"""ctypes binding of the C ABI declared in include/cagc_b200.h.

This is the only place the native library is loaded.  There is no fallback: if
`lib/libcagc_b200.so` is missing or its ABI version differs, importing this
module raises, and every op of the package is unusable ("no CPU
fallback in product code").  Build it with `python __graft_entry__.py` (or
`python -m b200gan.build`).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# CAGC_LIB: another build of the same library (A/B timing of two revisions on one box, scripts/build_rev.sh)
LIB_PATH = os.environ.get('CAGC_LIB') or os.path.join(os.path.dirname(_HERE), 'lib', 'libcagc_b200.so')
ABI_VERSION = 24

_p = C.c_void_p
_i = C.c_int
_l = C.c_int64
_f = C.c_float

# name -> (restype, argtypes); mirrors include/cagc_b200.h one to one
SIGNATURES = {
    'cagc_abi_version': (_i, []),
    'cagc_last_error': (C.c_char_p, []),
    'cagc_launch_count': (_l, []),
    'cagc_tc_available': (_i, []),
    'cagc_upfirdn2d': (_i, [_p, _p, _p, _p, _l, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    'cagc_fused_bias_act': (_i, [_p, _p, _p, _p, _p, _l, _l, _i, _i, _i, _f, _f]),
    'cagc_bias_grad_chunks': (_i, [_l]),
    'cagc_fused_bias_act_bwd': (_i, [_p, _p, _p, _p, _p, _l, _i, _l, _f, _f]),
    'cagc_bias_grad_rows_chunks': (_i, [_l, _i]),
    'cagc_fused_bias_act_bwd_rows': (_i, [_p, _p, _p, _p, _p, _l, _i, _f, _f]),
    'cagc_conv_same': (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _l, _i, _i]),
    'cagc_modulate': (_i, [_p, _p, _p, _p, _i, _i, _i, _i]),
    'cagc_conv_up': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i]),
    'cagc_conv_up_ws': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _l]),
    'cagc_conv_up_dgrad': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i]),
    'cagc_conv_wgrad_splits': (_i, [_i, _i, _i, _i, _i, _i, _i]),
    'cagc_conv_wgrad': (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
    'cagc_conv_wgrad_partial': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_int)]),
    'cagc_weight_prep': (_i, [_p, _p, _f, _i, _i, _i, _p, _i, _i, _i, _p, _i, _i, _i, _i, _p, _p, _i, _i, _p, _p, _i, _i]),
    'cagc_style_affine': (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _l, _i, _i, _i]),
    'cagc_style_affine_bwd': (_i, [_p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _l, _l, _p, _i, _i, _i]),
    'cagc_demod': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _f]),
    'cagc_act_bwd_finalize_blocks': (_i, [_i]),
    'cagc_act_bwd_finalize': (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i]),
    'cagc_style_grad_finalize': (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i]),
    'cagc_wgrad_finalize': (_i, [_p, _p, _i, _p, _f, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    'cagc_torgb_bwd_finalize': (_i, [_p, _p, _p, _p, _f, _p, _p, _i, _i, _i, _i, _i]),
    'cagc_linear_bias_act': (_i, [_p, _p, _p, _p, _i, _i, _f, _f, _i, _f, _f]),
    'cagc_linear_bias_act_bwd': (_i, [_p, _p, _p, _p, _p, _i, _i, _f, _f, _i, _f, _f]),
    'cagc_fir_nhwc': (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _l, _i]),
    'cagc_fir_nhwc_taps': (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _l, _i]),
    'cagc_act_bwd_chunks': (_i, [_i, _i]),
    'cagc_act_bwd': (_i, [_p, _p, _l, _l, _l, _l, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _l, _i]),
    'cagc_mod_bwd': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i]),
    'cagc_torgb_fwd': (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _i, _i, _i]),
    'cagc_torgb_bwd': (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f]),
    'cagc_to_nhwc': (_i, [_p, _p, _l, _l, _l, _l, _p, _i, _i, _i, _i, _i]),
    'cagc_adam_step': (_i, [_p, _p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _f, _f, _p]),
    'cagc_adam_ema_step': (_i, [_p, _p, _p, _p, _p, _l, _f, _f, _f, _f, _f, _f, _f, _p, _p, _f]),
    'cagc_conv_workspace_bytes': (_l, [_i, _i, _i, _i]),
    'cagc_conv_same_ws': (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _l, _i, _i, _p, _l]),
    'cagc_conv_same_psw_bytes': (_l, [_i, _i, _i, _i, _i, _i]),
    'cagc_conv_same_psw': (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _l, _i, _p, _l]),
    'cagc_conv_up_dgrad_ws': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _l]),
    'cagc_conv2d_ws': (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i, _p, _l]),
    'cagc_conv2d_mask_ws': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _l]),
    'cagc_linear_fwd': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _f, _f, _i, _f, _f]),
    'cagc_linear_bwd': (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i]),
    'cagc_conv2d': (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _i]),
    'cagc_fir_resample_nhwc': (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i]),
    'cagc_from_rgb_fwd': (_i, [_p, _p, _l, _l, _l, _l, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _f]),
    'cagc_from_rgb_bwd': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _i, _f]),
    'cagc_act_mask_nhwc': (_i, [_p, _p, _p, _p, _l, _f]),
    'cagc_rgb_conv3x3_fwd': (_i, [_p, _p, _l, _l, _l, _l, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i]),
    'cagc_rgb_conv3x3_bwd': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i]),
    'cagc_maxpool2_nhwc': (_i, [_p, _p, _p, _i, _i, _i, _i]),
    'cagc_relu_pool_bwd': (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i]),
    'cagc_lpips_head_blocks': (_i, [_i, _i]),
    'cagc_lpips_head_fwd': (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i]),
    'cagc_lpips_head_bwd': (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i]),
    'cagc_parse_preprocess': (_i, [_p, _p, _l, _l, _l, _l, _p, _i, _i, _i, _i]),
    'cagc_parsing_mask': (_i, [_p, _p, _p, _i, _i, _i, _i]),
    'cagc_parsing_mask_lowres': (_i, [_p, _p, _l, _l, _l, _l, _p, _i, _i, _i, _i, _i, _i]),
    'cagc_fir_nhwc_mask': (_i, [_p, _p, _p, _p, _p, _f, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f'native library {LIB_PATH} not found: build it with `python __graft_entry__.py` '
            '(nvcc -gencode arch=compute_100a,code=sm_100a); there is no fallback path')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    got = lib.cagc_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f'{LIB_PATH}: ABI version {got}, expected {ABI_VERSION}; rebuild the library')
    return lib


lib = _load()


class NativeError(RuntimeError):
    pass


def check(rc: int, what: str = ''):
    """Turn a non-zero return code into a Python exception (the reference raises RuntimeError
    through TORCH_CHECK, op/fused_bias_act.cpp:7-15)."""
    if rc != 0:
        msg = lib.cagc_last_error().decode('utf-8', 'replace')
        raise NativeError(f'{what or "cagc"} failed (code {rc}): {msg}')


def ptr(t):
    """Device pointer of a tensor or NULL for None."""
    return None if t is None else t.data_ptr()


def stream_of(t: torch.Tensor):
    """Current CUDA stream of the tensor's device (the reference launches on the current stream
    without a device guard, op/upfirdn2d_kernel.cu:213-215; callers here hold a device guard)."""
    return torch.cuda.current_stream(t.device).cuda_stream


def require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise RuntimeError(f'{what}: this implementation is CUDA-only (sm_100a); got a {t.device.type} tensor. '
                           'The CPU restatement lives in oracle/ and is test infrastructure.')
    if t.dtype != torch.float32:
        raise RuntimeError(f'{what}: only float32 tensors are supported, got {t.dtype}')


def conv_workspace(b: int, ho: int, wo: int, pitch: int, device):
    """Split-K scratch for a small convolution (None when the shape never splits).  Allocated per call from torch's
    caching allocator: stream-ordered, safe under CUDA-graph capture and with the teacher on a parallel branch."""
    n = int(lib.cagc_conv_workspace_bytes(b, ho, wo, pitch))
    if n <= 0:
        return None, 0
    return torch.empty(n // 4, device=device, dtype=torch.float32), n


def launch_count() -> int:
    return int(lib.cagc_launch_count())


class TensorCache:
    """Values derived from long-lived tensors (FIR buffers, frozen parameters), keyed by tensor IDENTITY and
    version counter.  A data pointer is not a safe key: the caching allocator hands the address of a freed
    temporary to the next tensor of the same size.  Entries die with their tensor (weak references)."""

    def __init__(self):
        self._d = {}

    def get(self, t: torch.Tensor, tag, make):
        key = (id(t), tag)
        hit = self._d.get(key)
        # identity + version counter + storage address: `module.to(device)` swaps `param.data` without touching
        # either the object or its version
        if hit is not None and hit[0]() is t and hit[1] == t._version and hit[3] == t.data_ptr():
            return hit[2]
        val = make()
        if val is None:
            return None
        d = self._d
        if len(d) > 8192:
            d.clear()
        d[key] = (weakref.ref(t, lambda _r, k=key, d=d: d.pop(k, None)), t._version, val, t.data_ptr())
        return val


_taps_cache = TensorCache()


def host_taps(fir: torch.Tensor):
    """Host copy (ctypes float array) of a small constant FIR buffer, cached per tensor and version.  Returns
    None while a CUDA graph is being captured and the buffer has not been seen before (a device->host copy is
    not capturable): callers then use the entry point that reads the taps from device memory."""
    def make():
        if torch.cuda.is_current_stream_capturing():
            return None
        vals = fir.detach().float().cpu().reshape(-1).tolist()
        return (C.c_float * len(vals))(*vals)
    return _taps_cache.get(fir, 'taps', make)


def fir_nhwc(st, x_ptr, fir: torch.Tensor, d, noise, nw, bias, out_ptr, b, h, w, pitch, valid, pads, nstride, act, what):
    """cagc_fir_nhwc / cagc_fir_nhwc_taps on raw NHWC-p buffers (pads = (x0, x1, y0, y1))."""
    kh, kw = fir.shape
    taps = host_taps(fir) if kh * kw <= 16 else None
    if taps is not None:
        check(lib.cagc_fir_nhwc_taps(st, x_ptr, fir.data_ptr(), taps, ptr(d), ptr(noise), ptr(nw), ptr(bias), out_ptr,
                                     b, h, w, pitch, valid, kh, kw, pads[0], pads[1], pads[2], pads[3], nstride, act), what)
    else:
        check(lib.cagc_fir_nhwc(st, x_ptr, fir.data_ptr(), ptr(d), ptr(noise), ptr(nw), ptr(bias), out_ptr,
                                b, h, w, pitch, valid, kh, kw, pads[0], pads[1], pads[2], pads[3], nstride, act), what)
