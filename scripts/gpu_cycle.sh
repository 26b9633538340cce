set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu --no-faithful > gpurun_out/bench_r2c_cfg3.json 2> gpurun_out/bench_r2c_cfg3.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-faithful > gpurun_out/bench_cycle.json 2> gpurun_out/bench_cycle.err
python -c "
import json
for f in ('bench_r2c_cfg3','bench_cycle'):
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1]);print(f, d['value'], d['stepwise']['value'], d['latency_ms_single_run'], d['e2e']['value'], d['e2e']['latency_ms_single_run'])"
