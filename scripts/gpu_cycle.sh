# one development cycle on the GPU box: the GPU parity suite, a short bench, the launch list of one run
set -x
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for b in 1 2; do
AOSLO_THIN_BLOCKS_PER_SM=$b timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-faithful > gpurun_out/bench_cycle_$b.json 2> gpurun_out/bench_cycle_$b.err
tail -3 gpurun_out/bench_cycle_$b.err
python -c "
import json;d=json.loads(open('gpurun_out/bench_cycle_$b.json').read().strip().splitlines()[-1]);print('blocks/SM $b', d['value'], d['latency_ms_single_run'], d['e2e'])"
done
AOSLO_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches_cycle.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-faithful --batch 1 > gpurun_out/ncu_launches_cycle.log 2>&1
python scripts/launch_table.py gpurun_out/launches_cycle.csv 2>/dev/null | head -12
