# ncu --set full capture of the fused DeepSets kernels of one bench configuration; prints the per-line summary
# usage: bash tools/ncu_deepsets.sh <config> <tag>
set -x
CFG=${1:-yaml}
TAG=${2:-ds}
ncu --set full --clock-control none --import-source on -k regex:"phi_pool_fwd|phi_bwd_chain|phi_wgrad_kernel" --launch-skip 6 -c 3 -o gpurun_out/ncu_$TAG python bench.py --config $CFG --steps 2 --warmup 3 --no-baselines > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_${TAG}_raw.csv 2>/dev/null
python tools/ncu_lines.py gpurun_out/ncu_$TAG.ncu-rep 30 > gpurun_out/ncu_${TAG}_lines.txt 2>&1
tail -2 gpurun_out/ncu_$TAG.log
