// tcgen05 / TMEM / TMA implicit-GEMM path for the modulated convolution (placeholder until the
// tensor-pipe kernel lands: reports "unsupported" so that callers fail loudly).
#include "conv_params.cuh"

int cagc_tc_conv(cudaStream_t, const cagc::ConvP&, const char* what) {
    return cagc::fail(CAGC_E_UNSUPPORTED, "%s: tcgen05 path not built", what);
}

extern "C" int cagc_tc_available(void) { return 0; }
