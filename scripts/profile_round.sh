set -x
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --batch 1 > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 1 --warmup 1 --no-cpu --batch 1 > gpurun_out/ncu_f.log 2>&1
timeout 200 python scripts/k1_bench.py --iters 3 --variants pair:auto,d:32 > gpurun_out/plain_k1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:blobness_pair -s 3 -c 1 -o gpurun_out/prof_k1_pair_f32 python scripts/k1_bench.py --iters 3 --variants pair:auto > gpurun_out/ncu_k1a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blobness_dtile -s 4 -c 1 -o gpurun_out/prof_k1_dtile_f64 python scripts/k1_bench.py --iters 3 --variants d:32 > gpurun_out/ncu_k1b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ns_force_move -s 45 -c 1 -o gpurun_out/prof_force python bench.py --steps 1 --warmup 1 --no-cpu --batch 1 > gpurun_out/ncu_force.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gravity_field -s 10 -c 1 -o gpurun_out/prof_gravity python bench.py --steps 1 --warmup 1 --no-cpu --batch 1 > gpurun_out/ncu_grav.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
