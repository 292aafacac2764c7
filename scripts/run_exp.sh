mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3) > gpurun_out/tests.log
(CAGC_KD_OVERLAP=0 timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench_noov.log
(timeout 200 python bench.py --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/bench.log
