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
