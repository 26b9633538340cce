set -x
for cfg in "512 1" "256 1" "256 2" "128 2"; do
set -- $cfg
AOSLO_THIN_THREADS=$1 AOSLO_THIN_BLOCKS_PER_SM=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-faithful > gpurun_out/bench_cycle_$1_$2.json 2> gpurun_out/bench_cycle.err
tail -3 gpurun_out/bench_cycle.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_cycle_$1_$2.json').read().strip().splitlines()[-1]);print('threads $1 blocks/SM $2:', d['value'], d['stepwise']['value'], d['latency_ms_single_run'], d['e2e']['value'], d['stepwise']['e2e'], d['e2e']['latency_ms_single_run'])"
done
