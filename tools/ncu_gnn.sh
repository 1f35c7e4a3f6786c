# ncu --set full capture of the fused GraphNet kernels (one launch each), raw + source pages exported next to it
set -x
TAG=${1:-r2d}
ncu --set full --clock-control none --import-source on -k regex:"gnn_conv_bwd|gnn_fc1_bwd|gnn_conv_fwd|gnn_agg_bwd" -c 4 -o gpurun_out/ncu_gnn_$TAG python tools/kt_graphnet_bf16.py > gpurun_out/ncu_gnn.log 2>&1
ncu -i gpurun_out/ncu_gnn_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_gnn_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_gnn_$TAG.ncu-rep --page source --csv --print-source cuda > gpurun_out/ncu_gnn_${TAG}_source.csv 2>/dev/null
tail -3 gpurun_out/ncu_gnn.log
