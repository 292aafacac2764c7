mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4) > gpurun_out/tests.log
(timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench.log
(TOP=70 timeout 300 python scripts/count_kernels.py 2>&1 | grep -E "style_affine|kernels/step|graph replay") > gpurun_out/census.log
