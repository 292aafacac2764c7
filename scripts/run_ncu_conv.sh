mkdir -p gpurun_out
CAGC_TC_HALO=2 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 8 -o gpurun_out/conv_halo python scripts/profile_layers.py student > gpurun_out/ncu_halo.log 2>&1
CAGC_TC_HALO=0 ncu --set full --clock-control none --import-source on -k regex:conv_tc -c 8 -o gpurun_out/conv_old python scripts/profile_layers.py student > gpurun_out/ncu_old.log 2>&1
