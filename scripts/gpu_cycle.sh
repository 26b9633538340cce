# last GPU call of the round: the parity suite on the final build, the default bench, one full ncu capture of the thinning kernel
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2c_final2.json 2> gpurun_out/bench_r2c_final2.err
tail -2 gpurun_out/bench_r2c_final2.err
AOSLO_NO_GRAPH=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:"thin_pixels" -c 1 -f -o gpurun_out/r02c_thin python bench.py --steps 1 --warmup 1 --no-cpu --no-faithful --batch 1 > gpurun_out/ncu_thin.log 2>&1
tail -2 gpurun_out/ncu_thin.log
