mkdir -p gpurun_out
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4) > gpurun_out/smoke.log
(timeout 600 python scripts/side_configs.py 2>&1 | tail -3) > gpurun_out/side.log
