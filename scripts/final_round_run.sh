# End-of-round measurements on one B200 (through gpurun; outputs land in gpurun_out/).
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2c_final.json 2> gpurun_out/bench_r2c_final.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2c_ref.json 2> gpurun_out/bench_r2c_ref.err
for w in cfg2 cfg3 sweep batch; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-faithful > gpurun_out/bench_r2c_$w.json 2> gpurun_out/bench_r2c_$w.err; done
timeout 200 python scripts/phase_cost.py > gpurun_out/phase_cost_final.log 2>&1
AOSLO_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-faithful --batch 1 > gpurun_out/ncu_launches_r02c.log 2>&1
AOSLO_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"thin_pixels|gravity_nearest|dual_metrics|occ_scatter|window_argmin|dual_rows|dual_force_move" --launch-skip 10 -c 9 -f -o gpurun_out/r02c_kernels python bench.py --steps 1 --warmup 1 --no-cpu --no-faithful --batch 1 > gpurun_out/ncu_kernels_r02c.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
