# compute-sanitizer memcheck over the tests of the non-tensor-pipe kernels (prep.cu, FIR, ToRGB, bias_act, SIMT convs)
mkdir -p gpurun_out
(timeout 700 compute-sanitizer --tool memcheck --print-limit 4 python -m pytest tests/test_prep_kernels.py tests/test_gpu_parity.py -q -x -k "prep or weight_prep or style_affine or finalizers or equal_linear or fir_nhwc or to_rgb_golden or generator_tiny_gradients or fused_act or upfirdn2d_golden or saliency" 2>&1 | grep -v "^$" | grep -E "=========|passed|failed|Error" | head -50) > gpurun_out/sanitizer.log
