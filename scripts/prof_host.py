"""Host-side profile of the end-to-end call: where the Python / driver time of find_blobs goes (cProfile, one lane,
2048^2 benchmark frame, canonical tie order, pinned host buffers)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench

lane = bench.Lane(0, torch.device("cuda", 0))
def e2e():
    lane.prep_e2e()          # find_blobs shifts the caller's GT in place: fresh copy per call
    return lane.run_e2e()


for _ in range(5):
    e2e()
torch.cuda.synchronize()
N = 50
t = time.perf_counter()
for _ in range(N):
    e2e()
torch.cuda.synchronize()
print("e2e wall per call: %.3f ms" % ((time.perf_counter() - t) / N * 1e3))
t = time.perf_counter()
for _ in range(N):
    lane.run_resident()
torch.cuda.synchronize()
print("resident wall per call: %.3f ms" % ((time.perf_counter() - t) / N * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    e2e()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
