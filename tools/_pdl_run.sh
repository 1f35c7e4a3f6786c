timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_pdl1.log 2> gpurun_out/bench_pdl1.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_pdl1.log').read().strip().splitlines()[-1]); print('PDL on ', d['value'], d['ms_per_step'], d['e2e']['value'])"
