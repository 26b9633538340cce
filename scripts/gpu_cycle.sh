# one development cycle on the GPU box: the GPU parity suite, a short bench, the launch list of one run
set -x
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-faithful > gpurun_out/bench_cycle.json 2> gpurun_out/bench_cycle.err
tail -3 gpurun_out/bench_cycle.err
AOSLO_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches_cycle.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-faithful --batch 1 > gpurun_out/ncu_launches_cycle.log 2>&1
python scripts/launch_table.py gpurun_out/launches_cycle.csv 2>/dev/null | head -42
