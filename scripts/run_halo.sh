mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3) > gpurun_out/tests.log
(CAGC_TC_HALO=0 timeout 120 python scripts/layer_times.py 2>&1 | tail -5) > gpurun_out/layers_h0.log
(CAGC_TC_HALO=1 timeout 120 python scripts/layer_times.py 2>&1 | tail -5) > gpurun_out/layers_h1.log
(timeout 200 python bench.py 2>&1 | tail -1) > gpurun_out/bench.log
