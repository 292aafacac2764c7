"""Drop-in replacement of the reference `op` package (op/__init__.py:1-2): same three names, backed
by the sm_100a kernels of libcagc_b200 through the C ABI (include/cagc_b200.h)."""
from .fused_act import FusedLeakyReLU, fused_leaky_relu
from .upfirdn2d import upfirdn2d

__all__ = ['FusedLeakyReLU', 'fused_leaky_relu', 'upfirdn2d']
