/*
 * cagc_b200.h -- C ABI of the B200-native StyleGAN2 generator hot path.
 *
 * This is the drop-in boundary.  The reference binds two
 * pybind11 modules at this level:
 *
 *   fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)   op/fused_bias_act.cpp:11-21
 *   upfirdn2d.upfirdn2d(input, kernel, up_x, up_y, down_x, down_y,
 *                       pad_x0, pad_x1, pad_y0, pad_y1)                  op/upfirdn2d.cpp:4-22
 *
 * and calls cuDNN/ATen for the modulated convolution (model.py:241-289).  Here
 * every one of those is a plain `extern "C"` entry point taking device
 * pointers, sizes and a CUDA stream: no torch types cross the boundary.  The
 * caller owns all memory (inputs, outputs, workspaces); nothing is retained;
 * every call is an asynchronous launch on `stream` and is re-entrant.
 *
 * Return value: 0 on success; a positive cudaError_t if a launch failed; a
 * negative CAGC_E_* code for rejected arguments.  `cagc_last_error()` returns
 * a thread-local human-readable message for the last non-zero return.
 *
 * Activation layout ("NHWC-p"): fp32, [B, H, W, pitch] with the channel pitch
 * a multiple of 4 floats (16 B; the host side uses multiples of 8).  Channels
 * >= the logical count are padding and are always written as zero.
 */
#ifndef CAGC_B200_H_
#define CAGC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cagc_stream_t; /* cudaStream_t */

#define CAGC_E_INVALID   (-1)  /* bad argument (null pointer, bad size, misaligned pitch) */
#define CAGC_E_UNSUPPORTED (-2) /* shape outside what the kernels implement */

/* bump when a signature changes; the Python loader checks it */
#define CAGC_ABI_VERSION 24

int cagc_abi_version(void);
const char* cagc_last_error(void);
/* number of kernel launches issued through this library by the calling process (all threads) */
int64_t cagc_launch_count(void);
/* 1 when the tcgen05/TMA tensor-pipe convolution path (algo 1) is compiled into this library */
int cagc_tc_available(void);

/* ------------------------------------------------------------------------
 * upfirdn2d -- replaces upfirdn2d.upfirdn2d (op/upfirdn2d.cpp:4-22,
 * op/upfirdn2d_kernel.cu:209-369).  Same tensor contract: input
 * [major, in_h, in_w, minor], kernel [kh, kw], output
 * [major, out_h, out_w, minor] with
 *   out_h = (in_h*up_y + pad_y0 + pad_y1 - kh) / down_y + 1   (op/upfirdn2d.py:103-104)
 * zero-insertion upsample -> pad/crop (negative pads crop) -> true 2-D
 * convolution -> decimation.  fp32 only.
 * ---------------------------------------------------------------------- */
int cagc_upfirdn2d(cagc_stream_t stream, const float* input, const float* kernel, float* output,
                   int64_t major, int in_h, int in_w, int minor, int kh, int kw,
                   int up_x, int up_y, int down_x, int down_y,
                   int pad_x0, int pad_x1, int pad_y0, int pad_y1);

/* ------------------------------------------------------------------------
 * fused_bias_act -- replaces fused.fused_bias_act (op/fused_bias_act.cpp:11-21,
 * op/fused_bias_act_kernel.cu:19-99).  x = in[i] (+ bias[(i / step_b) % size_b]
 * when bias != NULL); (act, grad) switch as in the reference: act 1 = linear,
 * 3 = leaky ReLU; grad 0: y = x>0 ? x : alpha*x; grad 1: y = refer>0 ? x :
 * alpha*x (refer = forward output); grad 2: y = 0.  out = y * scale.
 * ---------------------------------------------------------------------- */
int cagc_fused_bias_act(cagc_stream_t stream, const float* input, const float* bias, const float* refer,
                        float* output, int64_t n, int64_t step_b, int size_b,
                        int act, int grad, float alpha, float scale);

/* Backward of fused leaky ReLU with the bias gradient reduced in the same pass
 * (the reference runs a separate `grad_input.sum(dim)`, op/fused_act.py:33-39).
 * Layout [outer, size_b, step_b]; grad_in = (refer>0 ? g : alpha*g)*scale;
 * bias_partial is [outer * size_b * chunks] with chunks = cagc_bias_grad_chunks(step_b);
 * the caller sums it over (outer, chunks). */
int cagc_bias_grad_chunks(int64_t step_b);
int cagc_fused_bias_act_bwd(cagc_stream_t stream, const float* grad_out, const float* refer, float* grad_in,
                            float* bias_partial, int64_t outer, int size_b, int64_t step_b,
                            float alpha, float scale);

/* Same for channel-contiguous storage [rows][C] (channels-last 4-D tensors, 2-D [B, C] tensors):
 * bias_partial is [chunks][C], chunks = cagc_bias_grad_rows_chunks(rows, C) (0 = not eligible). */
int cagc_bias_grad_rows_chunks(int64_t rows, int C);
int cagc_fused_bias_act_bwd_rows(cagc_stream_t stream, const float* grad_out, const float* refer, float* grad_in,
                                 float* bias_partial, int64_t rows, int C, float alpha, float scale);

/* ------------------------------------------------------------------------
 * Modulated convolution (model.py:241-289), fused form of SURVEY.md App. B.
 *
 * cagc_conv_same: k x k stride-1 "same" convolution over NHWC-p input with
 * shared weights, optional per-(sample, in-channel) scale applied on load
 * (style modulation), and a fused epilogue
 *     y = out_scale[b,o] * u  (+ noise_w[0]*noise[b,y,x]) (+ bias[o]);  act: lrelu(0.2)*sqrt2
 * It is also the data-gradient kernel (pass the flipped / transposed weight
 * slabs).  Weight slabs: taps x [in_pitch rows][w_ld cols], zero padded,
 * tap t = ky*k + kx reads input pixel (y + ky - k/2, x + kx - k/2).
 * algo: 0 = fp32 SIMT (exact-order fp32, used by the saliency pass),
 *       1 = tcgen05 TF32 implicit GEMM (sm_100a tensor pipe).
 * ---------------------------------------------------------------------- */
int cagc_conv_same(cagc_stream_t stream, const float* in, const float* w_slabs, const float* in_scale,
                   const float* out_scale, const float* noise, const float* noise_w, const float* bias,
                   float* out, int B, int H, int W, int in_pitch, int out_pitch, int out_valid,
                   int ksize, int64_t noise_bstride, int act, int algo);

/* x~ = round_to_tf32(x * s[b, c]) on NHWC-p (s == NULL: rounding only).  The tcgen05 path (algo 1)
 * feeds the tensor pipe straight from shared memory filled by TMA, so it takes the style-modulated
 * input as a tensor: call this first and pass in_scale = NULL to cagc_conv_same / cagc_conv_up, with
 * K-major weight slabs taps x [roundup16(out_pitch) rows][in_pitch] instead of the SIMT layout. */
int cagc_modulate(cagc_stream_t stream, const float* x, const float* s, float* out, int B, int H, int W, int pitch);

/* Transposed stride-2 k x k convolution (model.py:259-267) of the modulated
 * input: out_T[b, 2y+ky, 2x+kx, o] += in[b,y,x,i]*in_scale[b,i]*W[t][i][o];
 * out_T is [B, 2H+k-2, 2W+k-2, out_pitch] and every element is written once
 * (4 phase convolutions, no atomics). */
int cagc_conv_up(cagc_stream_t stream, const float* in, const float* w_slabs, const float* in_scale,
                 float* out_t, int B, int H, int W, int in_pitch, int out_pitch, int ksize, int algo);

/* Data gradient of cagc_conv_up: g_in[b,y,x,i] = sum_{t,o} g_t[b,2y+ky,2x+kx,o] * Wd[t][o][i]
 * (weight slabs taps x [g_pitch rows][w_ld = in_pitch cols]). */
int cagc_conv_up_dgrad(cagc_stream_t stream, const float* g_t, const float* w_slabs, float* g_in,
                       int B, int H, int W, int g_pitch, int in_pitch, int ksize, int algo);

/* Weight gradient: gw[t][i][o] = sum_{b,y,x} a[b, y*sa+dya_t, x*sa+dxa_t, i]*a_scale[b,i] * g[b, y*sg+dyg_t, x*sg+dxg_t, o]
 *   mode 0 (same conv): base = output pixel, a = layer input at base + (ky-k/2, kx-k/2), g at base
 *   mode 1 (up conv)  : base = input pixel,  a at base, g = g_T at 2*base + (ky, kx)
 * Result slabs taps x [a_pitch][g_pitch]; `partial` is a workspace of
 * cagc_conv_wgrad_splits(...) such slab sets, reduced in a fixed order
 * (deterministic) into gw.  algo 1 (tcgen05, TF32): `a` must already be modulated (a_scale = NULL),
 * pitches >= 32 and g_pitch <= 256. */
int cagc_conv_wgrad_splits(int B, int H, int W, int a_pitch, int g_pitch, int ksize, int algo);
int cagc_conv_wgrad(cagc_stream_t stream, const float* a, const float* a_scale, const float* g,
                    float* gw, float* partial, int nsplits, int B, int H, int W,
                    int a_pitch, int g_pitch, int ksize, int mode, int algo);

/* Same as cagc_conv_wgrad without the final reduction: the `*nsplits_used` partial slab sets
 * [split][tap][a_pitch][g_pitch] are left in `partial` for cagc_wgrad_finalize. */
int cagc_conv_wgrad_partial(cagc_stream_t stream, const float* a, const float* a_scale, const float* g,
                            float* partial, int nsplits, int B, int H, int W,
                            int a_pitch, int g_pitch, int ksize, int mode, int algo, int* nsplits_used);

/* ------------------------------------------------------------------------
 * Layer bookkeeping around the convolution, one launch each (prep.cu).  They
 * replace the ATen elementwise / reduction / tiny-GEMM chains of
 * model.py:248-257 (weight scaling, modulation, demodulation) and of their
 * autograd backward.
 * ---------------------------------------------------------------------- */

/* W [O][I][k][k] (the reference parameter layout, model.py:225-227) -> operand slabs
 *   outA [k*k][RA][CA], element (t, o, i) = wscale*W[o][i][flipA ? k*k-1-t : t]   (nullable)
 *   outB [k*k][RB][CB], element (t, i, o) = wscale*W[o][i][flipB ? k*k-1-t : t]   (nullable)
 * zero padded, rounded to TF32 when round_tf32 != 0;
 *   wsq_oi [SO][SI], wsq_io [SI][SO] = sum_t (wscale*W)^2 (nullable, zero padded)   (model.py:252)
 *   bias_out [bias_np] = bias_in zero padded (nullable). */
int cagc_weight_prep(cagc_stream_t stream, const float* w, float wscale, int O, int I, int ksize,
                     float* outA, int RA, int CA, int flipA, float* outB, int RB, int CB, int flipB,
                     int round_tf32, float* wsq_oi, float* wsq_io, int SO, int SI,
                     const float* bias_in, float* bias_out, int bias_n, int bias_np);

/* Style modulation of up to 40 layers in one launch (EqualLinear, model.py:156-166 as used at :248):
 *   out_l[b][i] = scale_l * <A_l[i][:], latent[b][lat_l][:]> + bias_mul_l*bias_l[i]   (i < I_l), 0 for I_l <= i < pin_l
 * A, bias, out, I, pin, lat, scale, bias_mul are HOST arrays of n_layers entries (device pointers
 * inside); latent is [B][n_latent][D] with element strides (lat_sb, lat_sl, 1). */
int cagc_style_affine(cagc_stream_t stream, int n_layers, const void* const* A, const void* const* bias,
                      void* const* out, const int* I, const int* pin, const int* lat, const float* scale,
                      const float* bias_mul, const float* latent, int64_t lat_sb, int64_t lat_sl,
                      int B, int D, int n_latent);
/* Backward: gs_l [B][pin_l] (a NULL entry = zero gradient);
 *   gA_l[i][:] = scale_l * sum_b gs_l[b][i]*latent[b][lat_l][:],  gbias_l[i] = bias_mul_l * sum_b gs_l[b][i]
 *   g_latent[b][idx][:] = sum_{l: lat_l == idx} scale_l * sum_i gs_l[b][i]*A_l[i][:]   (contiguous [B][n_latent][D])
 * gA/gbias (host arrays) or g_latent may be NULL to skip that half. */
int cagc_style_affine_bwd(cagc_stream_t stream, int n_layers, const void* const* A, const void* const* gs,
                          void* const* gA, void* const* gbias, const int* I, const int* pin, const int* lat,
                          const float* scale, const float* bias_mul, const float* latent, int64_t lat_sb,
                          int64_t lat_sl, float* g_latent, int B, int D, int n_latent);

/* d[b][o] = rsqrt(sum_i s[b][i]^2 * wsq_io[i][o] + eps) for o < O, 0 for O <= o < pout (model.py:251-253) */
int cagc_demod(cagc_stream_t stream, const float* s, const float* wsq_io, float* d, int B, int pin, int O, int pout,
               float eps);

/* Reduce the chunk partials of cagc_act_bwd: g_bias[c] = sum_{b,chunk} partial[..][0][c];
 * gq[b][c] = -1/2 d[b][c]^3 * sum_chunk partial[b][..][1][c] (gradient w.r.t. the demodulation
 * radicand); nw_part[blk] = sum over the block's 32 channels of partial[..][2][..]
 * (blk < cagc_act_bwd_finalize_blocks(pitch); the caller adds them up).  Any output may be NULL. */
int cagc_act_bwd_finalize_blocks(int pitch);
int cagc_act_bwd_finalize(cagc_stream_t stream, const float* partial, const float* d, float* g_bias, float* gq,
                          float* nw_part, int B, int chunks, int pitch);

/* g_s[b][i] = sum_chunk mpartial[b][chunk][i] + 2*s[b][i]*sum_o gq[b][o]*wsq_oi[o][i]   (gq NULL: first term only) */
int cagc_style_grad_finalize(cagc_stream_t stream, const float* mpartial, const float* gq, const float* s,
                             const float* wsq_oi, float* g_s, int B, int chunks, int pin, int O, int pout);

/* W.grad in the parameter layout: out[o][i][t] = wscale * sum_split wpart[split][t][i][o]
 *                                              + 2*wscale^2*W[o][i][t] * sum_b gq[b][o]*s[b][i]^2   (gq NULL: first term) */
int cagc_wgrad_finalize(cagc_stream_t stream, const float* wpart, int nsplits, const float* w, float wscale,
                        const float* gq, const float* s, int B, int O, int I, int ksize, int pin, int pout,
                        float* out);

/* ToRGB backward bookkeeping from the chunk partials of cagc_torgb_bwd (t[b][o][i] = sum_chunk partial):
 *   g_w[b][o][i] = wscale * t[b][o][i]*s[b][i]  (per-sample contributions [B][nout][cin]; the caller adds them over b);
 *   g_s[b][i] = wscale * sum_o t[b][o][i]*w[o][i]  (pad channels 0) */
int cagc_torgb_bwd_finalize(cagc_stream_t stream, const float* partial, const float* s, const float* w,
                            float wscale, float* g_w, float* g_s, int B, int chunks, int cin, int pin, int nout);

/* EqualLinear epilogue (model.py:156-166): out = acc*acc_scale + bias*bias_scale, then (act != 0)
 * leaky ReLU(alpha) * gain; acc/out [M][N].  Backward: g_acc = g*act'*acc_scale, g_bias[n] = bias_scale*sum_m g*act'. */
int cagc_linear_bias_act(cagc_stream_t stream, const float* acc, const float* bias, float* out, int M, int N,
                         float acc_scale, float bias_scale, int act, float alpha, float gain);
int cagc_linear_bias_act_bwd(cagc_stream_t stream, const float* g, const float* out, float* g_acc, float* g_bias,
                             int M, int N, float acc_scale, float bias_scale, int act, float alpha, float gain);

/* NHWC-p FIR filter (up=down=1 case of upfirdn2d, used for the Blur after the
 * transposed conv, model.py:270) with the StyledConv epilogue fused:
 *   out = act( out_scale[b,c]*FIR(in) + noise_w*noise + bias[c] ).
 * fir is the kh x kw kernel as the caller would hand it to upfirdn2d (the
 * function flips it: true convolution). */
int cagc_fir_nhwc(cagc_stream_t stream, const float* in, const float* fir, const float* out_scale,
                  const float* noise, const float* noise_w, const float* bias, float* out,
                  int B, int in_h, int in_w, int pitch, int valid, int kh, int kw,
                  int pad_x0, int pad_x1, int pad_y0, int pad_y1, int64_t noise_bstride, int act);

/* Same with the kh*kw taps ALSO given on the host (row-major, as in `fir`): they travel in the kernel
 * parameter block (uniform registers) instead of being re-read from device memory by every thread.
 * `fir` must hold the same values (the fallback kernels read it). */
int cagc_fir_nhwc_taps(cagc_stream_t stream, const float* in, const float* fir, const float* taps_host,
                       const float* out_scale, const float* noise, const float* noise_w, const float* bias,
                       float* out, int B, int in_h, int in_w, int pitch, int valid, int kh, int kw,
                       int pad_x0, int pad_x1, int pad_y0, int pad_y1, int64_t noise_bstride, int act);

/* Backward of the StyledConv epilogue a = lrelu(d*u + nw*noise + bias)*sqrt2 (or of the bare
 * demodulation y = d*u when act == 0) on NHWC-p tensors:
 *   gz = ga*sqrt2*(a>0 ? 1 : 0.2);  gu = d*gz  (written, NHWC-p)
 *   partial[b][chunk][0][c] = sum_p gz ; [1] = sum_p gz*y/d ; [2] = sum_p gz*noise
 * ga may have arbitrary element strides (sb, sc, sh, sw). */
int cagc_act_bwd_chunks(int H, int W);
int cagc_act_bwd(cagc_stream_t stream, const float* ga, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                 const float* a, const float* d, const float* noise, const float* noise_w, const float* bias,
                 float* gu, float* partial, int B, int H, int W, int pitch, int valid,
                 int64_t noise_bstride, int act);

/* gs partials and input gradient of the modulation x~ = s*x:
 *   partial[b][chunk][c] = sum_p gxt[b,p,c]*x[b,p,c];  gxt <- s[b,c]*gxt (in place) */
int cagc_mod_bwd(cagc_stream_t stream, float* gxt, const float* x, const float* s, float* partial,
                 int B, int H, int W, int pitch);

/* ToRGB (model.py:380-395): 1x1 modulated conv without demodulation to `nout` (<= 4)
 * channels + bias + upsampled skip, output NCHW [B, nout, H, W].
 *   weff[b][o][i] = wscale * w[o][i] * s[b][i]
 * skip is NCHW [B, nout, H/2, W/2] or NULL; fir (fh x fw, already multiplied by up^2)
 * with pads as Upsample computes them (model.py:46-51). */
int cagc_torgb_fwd(cagc_stream_t stream, const float* x, const float* w, const float* s, const float* bias,
                   const float* skip, const float* fir, float* out, int B, int H, int W, int pitch,
                   int cin, int nout, float wscale, int fh, int fw, int pad0, int pad1);
/* gx[b,p,i] = sum_o weff[b,o,i]*g[b,o,p] (+ gx_add[b,p,i] when given: the gradient that reached the same
 * activation through the next convolution, so autograd needs no separate accumulation pass) (NHWC-p, pad
 * channels zero);  partial[b][chunk][o][pitch] = sum_p g[b,o,p]*x[b,p,i]  (caller reduces over chunk) */
int cagc_torgb_bwd(cagc_stream_t stream, const float* g, const float* x, const float* w, const float* s,
                   const float* gx_add, float* gx, float* partial, int B, int H, int W, int pitch, int cin, int nout,
                   float wscale);

/* NCHW (arbitrary strides) <-> NHWC-p conversion; pad channels zeroed. */
int cagc_to_nhwc(cagc_stream_t stream, const float* src, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                 float* dst, int B, int C, int H, int W, int pitch);

/* ------------------------------------------------------------------------
 * Fused Adam over one flat fp32 bucket (the step after the data-parallel
 * gradient allreduce, Miscellaneous/distributed.py:57-66 + train.py:308):
 *   g = grad[i]*grad_scale; m = b1*m + (1-b1)*g; v = b2*v + (1-b2)*g*g;
 *   p -= lr * (m/bc1) / (sqrt(v/bc2) + eps)     (torch.optim.Adam semantics)
 * step_dev (optional, device pointer to the step count t as a float): when given, the bias
 * corrections are computed on the device as 1 - beta^t, so the launch can be replayed from a CUDA graph.
 * ---------------------------------------------------------------------- */
int cagc_adam_step(cagc_stream_t stream, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   int64_t n, float lr, float beta1, float beta2, float eps, float grad_scale,
                   float bias_corr1, float bias_corr2, const float* step_dev);
/* Same, with the generator's exponential moving average folded into the pass (reference train.py:124-129
 * `accumulate(g_ema, g, decay)`, called after every optimizer step at :398): ema (nullable, bucket of the same
 * layout) <- ema * ema_decay + (1 - ema_decay) * updated param. */
int cagc_adam_ema_step(cagc_stream_t stream, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                       int64_t n, float lr, float beta1, float beta2, float eps, float grad_scale,
                       float bias_corr1, float bias_corr2, const float* step_dev, float* ema, float ema_decay);

/* ----------------------------------------------------------------------
 * Discriminator path (reference model.py:670-798: ConvLayer / ResBlock / EqualConv2d).  The reference runs
 * F.conv2d (cuDNN) + upfirdn2d + fused_bias_act launches; here the convolutions run on the same engines as the
 * generator (shared weights, no modulation) with bias / leaky ReLU / residual add in the epilogue.
 *
 * cagc_conv2d: out = act(conv(in, W) + bias) [+ residual] on NHWC-p.
 *   mode 0: stride 1, zero padding ksize/2 (replaces F.conv2d at model.py:117-124 for ConvLayer without
 *           downsample; with flipped / transposed slabs it is also its data gradient)
 *   mode 1: stride 2, no padding, H = 2*Ho + ksize - 2 (the EqualConv2d after Blur, model.py:683-700); its data
 *           gradient is cagc_conv_up (transposed convolution)
 *   act != 0: leaky ReLU(0.2) * act_gain (FusedLeakyReLU, op/fused_act.py:104-119; act_gain 0 means sqrt2);
 *           act_gain < 0: plain ReLU * |act_gain| (nn.ReLU of the VGG16 layers, lpips/pretrained_networks.py:100-114)
 *   residual (optional, layout of out, must not alias it) is added after the activation (ResBlock, model.py:731-737)
 *   w_slabs: as written by cagc_weight_prep for the engine `algo` (0 SIMT fp32, 1 tcgen05 TF32)
 * cagc_fir_resample_nhwc: upfirdn2d (op/upfirdn2d.py:145-156) with a 4x4 kernel and (up, down) = (1, 2) or (2, 1) on
 *   NHWC-p: Blur evaluated only where the following stride-2 1x1 convolution reads it (ResBlock skip branch,
 *   model.py:723-728) and its adjoint.  taps_host: the 16 kernel taps (host memory, row major, unflipped).
 * cagc_from_rgb_fwd / _bwd: ConvLayer(3, C, 1) (model.py:756): y = lrelu(W*scale*img + b) * gain straight to NHWC-p
 *   (img through element strides sb, sc, sh, sw); backward: NCHW-contiguous image gradient from (g, y).
 * cagc_act_mask_nhwc: out = g * gain * (y > 0 ? 1 : 0.2)  (fused_bias_act_kernel.cu:43, act=3 grad=1, no bias)
 * ---------------------------------------------------------------------- */
int cagc_conv2d(cagc_stream_t stream, const float* in, const float* w_slabs, const float* bias, const float* residual,
                float* out, int B, int Hin, int Win, int in_pitch, int out_pitch, int out_valid, int ksize, int mode,
                int act, float act_gain, int algo);
int cagc_fir_resample_nhwc(cagc_stream_t stream, const float* in, const float* taps_host, float* out, int B, int in_h,
                           int in_w, int pitch, int up, int down, int pad0, int pad1);
int cagc_from_rgb_fwd(cagc_stream_t stream, const float* img, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                      const float* w, const float* bias, float* out, int B, int H, int W, int cin, int cout, int pitch,
                      float wscale, int act, float gain);
int cagc_from_rgb_bwd(cagc_stream_t stream, const float* g, const float* yact, const float* w, float* gimg, int B,
                      int H, int W, int cin, int cout, int pitch, float wscale, int act, float gain);
int cagc_act_mask_nhwc(cagc_stream_t stream, const float* g, const float* y, float* out, int64_t n, float gain);
/* Blur^T + FusedLeakyReLU backward in one pass: out = upfirdn2d(in, fir, pad) * (mask_ref > 0 ? gain : 0.2 * gain)
 * (op/upfirdn2d.py:111-116 followed by fused_bias_act_kernel.cu:43); 4x4 kernels, up = down = 1, pitch % 8 == 0.
 * Returns CAGC_E_UNSUPPORTED when the shape is not handled (caller then runs cagc_fir_nhwc + cagc_act_mask_nhwc). */
int cagc_fir_nhwc_mask(cagc_stream_t stream, const float* in, const float* fir, const float* taps_host,
                       const float* mask_ref, float mask_gain, float* out, int B, int in_h, int in_w, int pitch, int valid,
                       int kh, int kw, int pad_x0, int pad_x1, int pad_y0, int pad_y1);

/* ----------------------------------------------------------------------
 * Split-K for layers too small to fill the machine (4x4 .. 16x16 images, or a per-rank batch of 2 under strong
 * scaling): the `_ws` forms of the convolution entry points take a caller-owned scratch buffer
 * (cagc_conv_workspace_bytes; 0 = this shape never splits).  The tensor-pipe engine may then deal the (tap, channel
 * chunk) iterations of the K loop to up to 8 CTAs per tile and add the partial sums in a fixed order in a second
 * pass that also applies the epilogue (deterministic).  Same results contract as the plain forms.
 * ---------------------------------------------------------------------- */
int64_t cagc_conv_workspace_bytes(int B, int Ho, int Wo, int out_pitch);
int cagc_conv_same_ws(cagc_stream_t stream, const float* in, const float* w_slabs, const float* in_scale,
                      const float* out_scale, const float* noise, const float* noise_w, const float* bias, float* out,
                      int B, int H, int W, int in_pitch, int out_pitch, int out_valid, int ksize, int64_t noise_bstride,
                      int act, int algo, float* workspace, int64_t workspace_bytes);
/* Same-resolution modulated convolution with the modulation folded into PER-SAMPLE weight slabs instead of a pass over
 * the activation (reference model.py:248-257 formulates it exactly so; here the per-sample set is B x k^2 K-major TF32
 * slabs written by one small kernel, and the implicit-GEMM kernel picks slab b * k^2 + tap for the tiles of sample b):
 * out = act(out_scale * conv(in, (w * in_scale[b])) + noise + bias).  Forward only (a layer that needs its weight
 * gradient keeps the modulated activation, which is the operand of that gradient).  tcgen05 engine only;
 * cagc_conv_same_psw_bytes: size of the w_ps scratch, 0 when the shape is not taken (the caller then modulates the
 * activation: cagc_modulate + cagc_conv_same_ws). */
int64_t cagc_conv_same_psw_bytes(int B, int H, int W, int in_pitch, int out_pitch, int ksize);
int cagc_conv_same_psw(cagc_stream_t stream, const float* in, const float* w_slabs, const float* in_scale,
                       const float* out_scale, const float* noise, const float* noise_w, const float* bias, float* out,
                       int B, int H, int W, int in_pitch, int out_pitch, int out_valid, int ksize, int64_t noise_bstride,
                       int act, float* w_ps, int64_t w_ps_bytes);
int cagc_conv_up_dgrad_ws(cagc_stream_t stream, const float* g_t, const float* w_slabs, float* g_in, int B, int H, int W,
                          int g_pitch, int in_pitch, int ksize, int algo, float* workspace, int64_t workspace_bytes);
/* cagc_conv_up with a workspace sized for the TRANSPOSED output (cagc_conv_workspace_bytes(B, 2H+k-2, 2W+k-2, out_pitch)):
 * low-resolution layers run all four phases and a split of their K loops as one launch plus a fixed-order reduction. */
int cagc_conv_up_ws(cagc_stream_t stream, const float* in, const float* w_slabs, const float* in_scale, float* out_t,
                    int B, int H, int W, int in_pitch, int out_pitch, int ksize, int algo, float* workspace,
                    int64_t workspace_bytes);
int cagc_conv2d_ws(cagc_stream_t stream, const float* in, const float* w_slabs, const float* bias, const float* residual,
                   float* out, int B, int Hin, int Win, int in_pitch, int out_pitch, int out_valid, int ksize, int mode,
                   int act, float act_gain, int algo, float* workspace, int64_t workspace_bytes);
/* out = conv(in, W) * (mask_ref > 0), stride 1, same size: a data-gradient convolution (flipped / transposed slabs) with
 * the backward of the ReLU that produced its input activation fused into the epilogue (VGG16 of the LPIPS loss:
 * conv -> ReLU -> conv chains, lpips/pretrained_networks.py:100-114; replaces a threshold_backward pass per layer).
 * mask_ref: the forward activation, layout of out. */
int cagc_conv2d_mask_ws(cagc_stream_t stream, const float* in, const float* w_slabs, const float* mask_ref, float* out,
                        int B, int Hin, int Win, int in_pitch, int out_pitch, int out_valid, int ksize, int algo,
                        float* workspace, int64_t workspace_bytes);

/* ----------------------------------------------------------------------
 * EqualLinear (reference model.py:137-166: F.linear(x, W * scale) + fused_leaky_relu(bias * lr_mul)) as one launch:
 *   out[M,N] = act(acc_scale * x[M,K] W[N,K]^T + bias[N] * bias_scale) (* gain when act != 0, leaky slope alpha)
 * cagc_linear_bwd: g_x[M,K] = g_acc[M,N] W (nullable), g_w[N,K] = g_acc^T x (nullable); g_acc is the output of
 * cagc_linear_bias_act_bwd (activation mask and acc_scale already applied).  Fixed summation order.
 * ---------------------------------------------------------------------- */
int cagc_linear_fwd(cagc_stream_t stream, const float* x, const float* w, const float* bias, float* out, int M, int N,
                    int K, float acc_scale, float bias_scale, int act, float alpha, float gain);
int cagc_linear_bwd(cagc_stream_t stream, const float* g_acc, const float* x, const float* w, float* g_x, float* g_w,
                    int M, int N, int K);

/* ----------------------------------------------------------------------
 * LPIPS-VGG16 perceptual distance of the KD loss (reference train.py:172-182 -> lpips/__init__.py:13-41 ->
 * lpips/networks_basic.py:26-92 -> lpips/pretrained_networks.py:97-137), frozen network, NHWC-p activations.
 * The 3x3 convolutions with >= 64 input channels are cagc_conv2d[_ws] calls with act_gain = -1 (ReLU); these are the
 * remaining pieces (replacing ATen / cuDNN launches of the reference: F.conv2d on 3 channels, max_pool2d and its
 * backward, threshold_backward, the normalize / square / 1x1-conv / mean chain of networks_basic.py:66-84).
 *   cagc_rgb_conv3x3_fwd: out[B,H,W,pitch] = relu(conv3x3((img - shift) / scale, w[cout,3,3,3], pad 1) + bias); img through
 *       element strides (sb, sc, sh, sw); shift_host / scale_host: 3 floats in HOST memory (ScalingLayer,
 *       networks_basic.py:94-101; null = identity).  _bwd: NCHW-contiguous image gradient from gz[B,H,W,pitch] (the
 *       gradient w.r.t. the pre-activation, i.e. already masked).
 *   cagc_maxpool2_nhwc: nn.MaxPool2d(kernel_size=2, stride=2); H, W even.
 *   cagc_relu_pool_bwd: out = ((act wins its 2x2 window ? g_pool[B,H/2,W/2,pitch] : 0) + g_direct) * (act > 0); the first
 *       maximum in row-major order wins (ATen).  g_pool null: out = g_direct * (act > 0).  g_direct may be null.
 *   cagc_lpips_head_fwd: val[b] (+)= mean_p sum_c lin_w[c] (fs/(|fs|+1e-10) - ft/(|ft|+1e-10))^2 for feature maps
 *       [B,HW,C] (C in {64,128,256,512}); partial: scratch of B * cagc_lpips_head_blocks(B, HW) floats; fixed summation
 *       order.  _bwd: gs = gval[b] * d val[b] / d fs.
 * ---------------------------------------------------------------------- */
int cagc_rgb_conv3x3_fwd(cagc_stream_t stream, const float* img, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                         const float* w, const float* bias, const float* shift_host, const float* scale_host, float* out,
                         int B, int H, int W, int cout, int pitch);
int cagc_rgb_conv3x3_bwd(cagc_stream_t stream, const float* gz, const float* w, const float* scale_host, float* gimg, int B,
                         int H, int W, int cout, int pitch);
int cagc_maxpool2_nhwc(cagc_stream_t stream, const float* in, float* out, int B, int H, int W, int pitch);
int cagc_relu_pool_bwd(cagc_stream_t stream, const float* act, const float* g_pool, const float* g_direct, float* out,
                       int B, int H, int W, int pitch);
int cagc_lpips_head_blocks(int B, int HW);
int cagc_lpips_head_fwd(cagc_stream_t stream, const float* fs, const float* ft, const float* lin_w, float* partial,
                        float* val, int B, int HW, int C, int accumulate);
int cagc_lpips_head_bwd(cagc_stream_t stream, const float* fs, const float* ft, const float* lin_w, const float* gval,
                        float* gs, int B, int HW, int C);

/* ----------------------------------------------------------------------
 * Content-mask glue of the KD loss (reference Util/content_aware_pruning.py:61-117, train.py:154-158).
 *   cagc_parse_preprocess: out[N,3,P,P] (contiguous, or stored [N,P,P,3] when channels_last != 0) = (bilinear(clamp((img + 1) / 2, 0, 1), S -> P) - mean) / std with
 *       the ImageNet statistics of Batch_Img_Parsing (:71-82); img [N,3,S,S] through element strides.
 *   cagc_parsing_mask: mask[N,S,S] = bilinear(float(argmax_k logits[N,K,P,P] not in {0, 16}), P -> S) > 0.5
 *       (`Batch_Img_Parsing` :87 + `Get_Masked_Tensor` :103-109, without the host round trip); K == 1: the single plane
 *       holds the labels themselves (as floats).
 * Bilinear = F.interpolate(align_corners=False, scale_factor = out / in).
 * ---------------------------------------------------------------------- */
int cagc_parse_preprocess(cagc_stream_t stream, const float* img, int64_t sb, int64_t sc, int64_t sh, int64_t sw,
                          float* out, int N, int S, int P, int channels_last);
int cagc_parsing_mask(cagc_stream_t stream, const float* logits, float* mask, int N, int K, int P, int S);
/* The same mask from the parser's low-resolution scores [N,K,h,w] (element strides sn, sk, sh, sw): the final
 * F.interpolate(scores, (P, P), bilinear, align_corners=True) of BiSeNet.forward (Util/face_parsing/BiSeNet.py:247) is
 * evaluated inside the kernel instead of being materialised. */
int cagc_parsing_mask_lowres(cagc_stream_t stream, const float* scores, int64_t sn, int64_t sk, int64_t sh, int64_t sw,
                             float* mask, int N, int K, int h, int w, int P, int S);

#ifdef __cplusplus
}
#endif
#endif /* CAGC_B200_H_ */
