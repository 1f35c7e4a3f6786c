#!/bin/bash
# usage: tools/run_multi.sh N tag "configs..."   (one bench.py line per config at N GPUs -> gpurun_out/scale_<config>_<N>_<tag>.json)
N=$1; TAG=$2; shift 2
for c in "$@"; do
  steps=20; [ "$c" = "graphnet" ] && steps=10; [ "$c" = "sweep" ] && steps=10
  if [ "$N" = "1" ]; then
    python bench.py --config $c --gpus 1 --steps $steps --warmup 5 --no-baselines > gpurun_out/scale_${c}_${N}_${TAG}.json 2> gpurun_out/scale_${c}_${N}_${TAG}.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 200)) \
      bench.py --config $c --gpus $N --steps $steps --warmup 5 --no-baselines > gpurun_out/scale_${c}_${N}_${TAG}.json 2> gpurun_out/scale_${c}_${N}_${TAG}.err
  fi
  echo "== $c N=$N rc=$?"; cut -c1-260 gpurun_out/scale_${c}_${N}_${TAG}.json; tail -2 gpurun_out/scale_${c}_${N}_${TAG}.err
done
