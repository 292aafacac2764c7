# Round-end evidence (1 GPU): bench line, ncu launch list of one KD step (+DRAM bytes), ncu --set full of the layer kernels.
mkdir -p gpurun_out
(timeout 300 python bench.py 2>/dev/null | tail -1) > gpurun_out/r1_bench_1gpu.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/r1_launches_kdstep.csv python scripts/profile_step.py > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'conv_tc|wgrad_tc|fir_nhwc' -c 60 -o gpurun_out/r1_layers python scripts/profile_layers.py > gpurun_out/layers_ncu.log 2>&1
ncu -i gpurun_out/r1_layers.ncu-rep --page raw --csv > gpurun_out/r1_ncu_full_layers_raw.csv 2>/dev/null
ls -la gpurun_out | head -30
