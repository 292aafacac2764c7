"""Process-wide switches of the hot path (thread-local so DataParallel replicas do not race)."""
import contextlib
import threading

_state = threading.local()

ALGO_SIMT_FP32 = 0   # exact fp32 SIMT engine (saliency pass, validation)
ALGO_TCGEN05_TF32 = 1  # sm_100a tensor pipe, TF32 operands / fp32 accumulate (reference cuDNN convs also run TF32)

_default_algo = ALGO_SIMT_FP32


def set_default_algo(algo: int):
    global _default_algo
    _default_algo = int(algo)


def best_available_algo() -> int:
    from ._lib import lib
    return ALGO_TCGEN05_TF32 if lib.cagc_tc_available() else ALGO_SIMT_FP32


def conv_algo() -> int:
    return getattr(_state, 'algo', _default_algo)


@contextlib.contextmanager
def use_algo(algo: int):
    prev = getattr(_state, 'algo', None)
    _state.algo = int(algo)
    try:
        yield
    finally:
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
