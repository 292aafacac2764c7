"""Process-wide switches of the hot path.

Convolution engine selection, in order of precedence:
  1. the calling thread's `use_algo(...)` / `exact_fp32()` context,
  2. a context entered by ANOTHER thread (nn.DataParallel runs the replicas in worker threads while the caller
     sits inside its `with` block -- train.py:522-525, get_fid.py:27),
  3. the process default: `CAGC_CONV_ALGO` = `tc` | `simt` | `auto` (default `auto`: the tcgen05 TF32 path when
     the current device is sm_100, which is what an unmodified train.py / get_fid.py therefore runs on).
The autograd backward of a layer uses the engine its forward captured (it runs on an autograd thread).
"""
import contextlib
import os
import threading

_state = threading.local()

ALGO_SIMT_FP32 = 0   # exact fp32 SIMT engine (saliency pass, validation)
ALGO_TCGEN05_TF32 = 1  # sm_100a tensor pipe, TF32 operands / fp32 accumulate (reference cuDNN convs also run TF32)

_default_algo = {'simt': ALGO_SIMT_FP32, 'tc': ALGO_TCGEN05_TF32}.get(os.environ.get('CAGC_CONV_ALGO', 'auto').lower())
_shared_lock = threading.Lock()
_shared_stack = []      # contexts currently open in any thread (innermost last)


def set_default_algo(algo: int):
    global _default_algo
    _default_algo = int(algo)


def best_available_algo() -> int:
    from ._lib import lib
    return ALGO_TCGEN05_TF32 if lib.cagc_tc_available() else ALGO_SIMT_FP32


def conv_algo() -> int:
    global _default_algo
    a = getattr(_state, 'algo', None)
    if a is not None:
        return a
    if _shared_stack:
        try:
            return _shared_stack[-1]
        except IndexError:
            pass
    if _default_algo is None:
        _default_algo = best_available_algo()      # resolved on first use: needs the CUDA context
    return _default_algo


@contextlib.contextmanager
def use_algo(algo: int):
    prev = getattr(_state, 'algo', None)
    _state.algo = int(algo)
    with _shared_lock:
        _shared_stack.append(int(algo))
    try:
        yield
    finally:
        with _shared_lock:
            for i in range(len(_shared_stack) - 1, -1, -1):
                if _shared_stack[i] == int(algo):
                    del _shared_stack[i]
                    break
        if prev is None:
            del _state.algo
        else:
            _state.algo = prev


def exact_fp32():
    """Context for the content-aware saliency pass: fp32 SIMT convolutions, fixed reduction order."""
    return use_algo(ALGO_SIMT_FP32)


def is_second_order() -> bool:
    return getattr(_state, 'second_order', False)


@contextlib.contextmanager
def second_order():
    """Route the modulated convolutions through the differentiable composite (torch ops + our
    upfirdn2d / fused_leaky_relu, which implement double backward) -- used by the path-length
    regulariser (model.py:661-666), which needs create_graph=True."""
    prev = is_second_order()
    _state.second_order = True
    try:
        yield
    finally:
        _state.second_order = prev


# ------------------------------------------------------------------------------------------------
# optional per-kernel timing (bench.py): CUDA events around each native convolution / FIR launch
# ------------------------------------------------------------------------------------------------
_profiler = None


class KernelProfiler:
    """Collects (kernel name -> algorithmic work, CUDA-event pairs).  Events are recorded on the
    stream the kernels are launched on (torch's current stream)."""

    def __init__(self):
        self.records = {}

    def add(self, name, flops, nbytes, e0, e1):
        r = self.records.setdefault(name, {'flops': 0.0, 'bytes': 0.0, 'events': [], 'launches': 0})
        r['flops'] += flops
        r['bytes'] += nbytes
        r['events'].append((e0, e1))
        r['launches'] += 1

    def summary(self):
        out = {}
        for name, r in self.records.items():
            ms = sum(a.elapsed_time(b) for a, b in r['events'])
            out[name] = {'ms': ms, 'launches': r['launches'], 'flops': r['flops'], 'bytes': r['bytes']}
        return out


def set_profiler(p):
    global _profiler
    _profiler = p


def profiler():
    return _profiler
