mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -12) > gpurun_out/tests.log
(CAGC_TC_WGRAD_ALL=2 timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -12) > gpurun_out/tests_all2.log
(timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench.log
(CAGC_TC_WGRAD_ALL=0 timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_w0.log
(CAGC_TC_WGRAD_ALL=2 timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_w2.log
