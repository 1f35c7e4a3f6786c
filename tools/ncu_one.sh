# ncu --set full capture of kernels matching $1 (regex) in the GraphNet step, tag $2; prints the per-line summary
set -x
K=${1:-csrt_block}
TAG=${2:-one}
ncu --set full --clock-control none --import-source on -k regex:"$K" -c ${3:-2} -o gpurun_out/ncu_$TAG python tools/kt_graphnet_bf16.py > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv 2>/dev/null
python tools/ncu_lines.py gpurun_out/ncu_$TAG.ncu-rep 18 > gpurun_out/ncu_${TAG}_lines.txt 2>&1
tail -2 gpurun_out/ncu_$TAG.log
