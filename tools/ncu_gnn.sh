set -x
ncu --set full --clock-control none --import-source on -k regex:"gnn_agg_bwd|gnn_conv_fwd" -c 2 -o gpurun_out/ncu_gnn_r2a python tools/kt_graphnet_bf16.py > gpurun_out/ncu_gnn.log 2>&1
ncu -i gpurun_out/ncu_gnn_r2a.ncu-rep --page raw --csv > gpurun_out/ncu_gnn_r2a_raw.csv 2>/dev/null
tail -3 gpurun_out/ncu_gnn.log
