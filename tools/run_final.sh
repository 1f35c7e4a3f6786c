# Round-end measurement pass on one B200 (run under gpurun): GPU test suite, smoke, every bench configuration, the
# GraphNet kernel table, ncu launch lists and one ncu --set full capture of the fused GraphNet kernels.
# usage: bash tools/run_final.sh <tag>     (outputs: gpurun_out/final_<tag>/)
TAG=${1:-r2}
OUT=gpurun_out/final_$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.txt 2>&1; tail -3 $OUT/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $OUT/smoke.txt 2>&1; tail -1 $OUT/smoke.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
for c in yaml ragged graphnet; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 5 > $OUT/bench_$c.json 2> $OUT/bench_$c.err
done
timeout 900 python bench.py --config sweep --steps 10 --warmup 3 > $OUT/bench_sweep.json 2> $OUT/bench_sweep.err
timeout 300 python tools/kt_graphnet_bf16.py > $OUT/kt_gnn_bf16.txt 2>&1
timeout 300 python tools/bench_knn.py > $OUT/bench_knn.txt 2>&1
# launch lists (ncu serialises and cold-starts every kernel: shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-baselines > $OUT/ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_graphnet.csv python bench.py --config graphnet --steps 2 --warmup 1 --no-baselines > $OUT/ncu_launch_g.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gnn_conv_bwd|gnn_fc1_bwd|gnn_conv_fwd|gnn_agg_bwd|knn_tiled|csrt_block" -c 6 -o $OUT/ncu_gnn python tools/kt_graphnet_bf16.py > $OUT/ncu_gnn.log 2>&1
ncu -i $OUT/ncu_gnn.ncu-rep --page raw --csv > $OUT/ncu_gnn_raw.csv 2>/dev/null
python tools/ncu_lines.py $OUT/ncu_gnn.ncu-rep 12 > $OUT/ncu_gnn_lines.txt 2>&1
rm -f $OUT/ncu_gnn.ncu-rep
ls -la $OUT
