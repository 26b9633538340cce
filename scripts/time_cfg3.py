"""BASELINE config 3: one detection run on a synthetic 4096^2 montage (single GPU, tiled blobness with halos),
resident inputs, CUDA events; prints latency and the K1 share."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from detecting_cones_aoslo_b200 import _lib
from detecting_cones_aoslo_b200.utils import metrics_from_record

size = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bench.SIZE = size
bench.make_workload_orig = bench.make_workload
bench.make_workload = lambda rank, size=size: bench.make_workload_orig(rank, size)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
ln = bench.Lane(0, dev)
for _ in range(2):
    ln.run_resident()
torch.cuda.synchronize()
ms = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ln.stream):
        a.record()
    s = ln.run_resident(time_k1=True)
    with torch.cuda.stream(ln.stream):
        b.record()
    torch.cuda.synchronize()
    ms.append(a.elapsed_time(b))
k1 = [x.elapsed_time(y) for x, y in ln.k1_events]
summ, recs, _ = ln.last
pm, *_ = metrics_from_record(recs[-1])
print(json.dumps({"size": size, "padded": [ln.H, ln.W], "n_gt": ln.n_gt, "n_seeds": int(summ.n_seeds),
                  "n_particles": int(summ.n_particles), "iterations": int(summ.n_iterations),
                  "latency_ms": float(np.median(ms)), "k1_f64_ms": float(np.median(k1)),
                  "dice_2": float(pm['2']['dice'])}))
