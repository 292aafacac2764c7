mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4) > gpurun_out/tests.log
(timeout 400 python bench.py --size 1024 --batch 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200) > gpurun_out/bench1024.log
(timeout 300 python bench.py 2>&1 | tail -1) > gpurun_out/bench.log
