# compute-sanitizer memcheck over the GPU parity tests (SURVEY.md §5 "race detection / sanitizers"):
#   bash scripts/sanitize.sh            (on a B200 box; ~1 minute)
# set 1: prep.cu, streaming FIR (TMA), ToRGB, bias_act, SIMT convolutions, saliency; set 2: the tcgen05 kernels incl. full-size
mkdir -p gpurun_out
(timeout 700 compute-sanitizer --tool memcheck --print-limit 4 python -m pytest tests/test_prep_kernels.py tests/test_gpu_parity.py -q -x -k "prep or weight_prep or style_affine or finalizers or equal_linear or fir_nhwc or to_rgb_golden or generator_tiny_gradients or fused_act or upfirdn2d_golden or saliency" 2>&1 | grep -v "^$" | grep -E "=========|passed|failed|Error" | head -50) > gpurun_out/sanitizer.log
(timeout 400 compute-sanitizer --tool memcheck --print-limit 4 python -m pytest tests/test_gpu_parity.py -q -x -k "tcgen05_styled_conv_vs_oracle or tcgen05_modconv_gradients_strict or full_size" 2>&1 | grep -v "^$" | grep -E "=========|passed|failed|Error" | head -40) > gpurun_out/sanitizer_tc.log
# racecheck (shared-memory hazards) over the kernels that stage through shared memory without the tensor pipe
(timeout 400 compute-sanitizer --tool racecheck --print-limit 6 python -m pytest tests/test_prep_kernels.py tests/test_gpu_parity.py -q -x -k "finalizers or style_affine or fir_nhwc_streaming or to_rgb_golden or fused_act_golden" 2>&1 | grep -v "^$" | grep -E "=========|passed|failed|Error" | cut -c1-220 | head -40) > gpurun_out/racecheck.log
