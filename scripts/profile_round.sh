#!/bin/bash
# Profile artefacts of a round (run on the B200 box through gpurun; outputs land in gpurun_out/, summaries are
# written into profiles/ by the python snippets in scripts/summarise_profiles.py / scripts/launch_table.py).
# ncu cannot profile kernel nodes of graphs that contain conditional nodes, so the particle kernels are captured on
# the eager path of aoslo_run (AOSLO_NO_GRAPH=1: same kernels, loop conditions evaluated on the host).
set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-faithful --batch 1"
# 1. the program must exit 0 without ncu first
$B > gpurun_out/plain.log 2>&1 || exit 1
# 2. launch list of one run
AOSLO_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv \
    --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_launches.log 2>&1
# 3. full captures of the top kernels
AOSLO_NO_GRAPH=1 ncu --set full --clock-control none --import-source on \
    -k regex:"thin_pixels|dual_metrics|dual_force_move|dual_rows|gravity_nearest|occ_scatter|window_argmin|ns_relink|period_begin" \
    --launch-skip 0 -c 16 -f -o gpurun_out/r02_kernels $B > gpurun_out/ncu_kernels.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blobness_dtile|blobness_pair" --launch-skip 6 -c 2 -f \
    -o gpurun_out/r02_k1 python scripts/k1_bench.py --variants f64:32:2,f32:auto > gpurun_out/ncu_k1.log 2>&1
# 4. K1 variants, FP64 pipe and conditional-node probes
python scripts/k1_bench.py > gpurun_out/k1_r02.log 2>&1
./scripts/probes/fp64_probe > gpurun_out/fp64_probe.log 2>&1
./scripts/probes/cond_graph_probe > gpurun_out/cond_probe.log 2>&1
