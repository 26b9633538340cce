set -x
for cfg in "128 2" "256 2" "512 2" "256 4" "512 4"; do
set -- $cfg
AOSLO_THIN_THREADS=$1 timeout 300 python bench.py --workload cfg3 --batch $2 --steps 5 --warmup 3 --no-cpu --no-faithful > gpurun_out/bench_cfg3_$1_$2.json 2> gpurun_out/bench_cycle.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_cfg3_$1_$2.json').read().strip().splitlines()[-1]);print('thin threads $1 lanes $2:', d['value'], d['stepwise']['value'], d['latency_ms_single_run'], d['e2e']['value'], d['e2e']['latency_ms_single_run'])"
done
