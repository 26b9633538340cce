set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_final2.json 2> gpurun_out/bench_r2_final2.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2_ref2.json 2> gpurun_out/bench_r2_ref2.err
for w in cfg2 cfg3 sweep batch; do timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-faithful > gpurun_out/bench_r02b_$w.json 2> gpurun_out/bench_r02b_$w.err; done
timeout 300 python scripts/k1_bench.py > gpurun_out/k1_r02_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"blobness_dtile" --launch-skip 6 -c 1 -f -o gpurun_out/r02_k1_final python scripts/k1_bench.py --variants f64:0:0 > gpurun_out/ncu_k1_final.log 2>&1
AOSLO_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-faithful --batch 1 > gpurun_out/ncu_launches_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
