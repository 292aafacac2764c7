mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2) > gpurun_out/tests.log
(TOP=12 timeout 300 python scripts/count_kernels.py 2>&1 | tail -14) > gpurun_out/census.log
(timeout 200 python bench.py --no-cpu-baseline --steps 30 2>&1 | tail -1) > gpurun_out/bench.log
