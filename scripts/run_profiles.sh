# Round-end evidence (1 GPU): bench lines, ncu launch list of one complete KD step (+DRAM bytes), ncu --set full of the
# generator / discriminator layer kernels.  Keep what is written under gpurun_out/ below 64 MiB (it travels back):
# --set full over the LPIPS step with --import-source does not fit.
mkdir -p gpurun_out
(timeout 400 python bench.py 2>/dev/null | tail -1) > gpurun_out/r2_bench_1gpu.json
(timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1) > gpurun_out/r2_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/r2_launches_kdstep_b16.csv python scripts/profile_step.py > gpurun_out/launches.log 2>&1
ncu --set full --clock-control none --profile-from-start off \
    -k regex:'conv_tc|wgrad_tc|fir_nhwc' -c 40 -o gpurun_out/r2_layers python scripts/profile_layers.py > gpurun_out/layers_ncu.log 2>&1
ncu -i gpurun_out/r2_layers.ncu-rep --page raw --csv > gpurun_out/r2_layers_raw.csv 2>/dev/null
python scripts/ncu_compact.py gpurun_out/r2_layers_raw.csv > gpurun_out/r2_ncu_full_layers.csv
rm -f gpurun_out/r2_layers_raw.csv
ls -la gpurun_out | head -30
# then: copy the artefacts to profiles/ and run `python scripts/summarize_r2.py`
