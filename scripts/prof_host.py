import cProfile, pstats, sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench
from detecting_cones_aoslo_b200.pipeline_with_delitions import find_blobs
img, gt, cfg = bench.make_workload(0)
for _ in range(2):
    find_blobs(dict(cfg), gt.copy(), img, None, False, False)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
t=time.perf_counter()
for _ in range(3):
    find_blobs(dict(cfg), gt.copy(), img, None, False, False)
torch.cuda.synchronize()
print('per call ms', (time.perf_counter()-t)/3*1e3)
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(25)
