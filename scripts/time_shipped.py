"""Latency of the drop-in find_blobs on the shipped annotated image (config_zoo), both tie orders.  (The CPU
oracle is test infrastructure: its timing on the same inputs comes from bench.py's cpu_baseline leg and tests/.)"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from detecting_cones_aoslo_b200 import configs
from detecting_cones_aoslo_b200.pipeline_with_delitions import find_blobs
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dataset.npz"))
img, gt = d["img"], d["gt"].astype(np.float64)
out = {}
for tie in ("sklearn", "canonical"):
    cfg = configs.variant(configs.zoo_config(), tie_mode=tie)
    for _ in range(3):
        res, _ = find_blobs(dict(cfg), gt.copy(), img, None, False, False)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        res, ts = find_blobs(dict(cfg), gt.copy(), img, None, False, False)
    torch.cuda.synchronize()
    out[tie] = {"ms_per_run": (time.perf_counter() - t) / 10 * 1e3, "dice_2": res["dice_coeff_2"],
                "n": res["_summary"]["n_particles"], "iters": res["_summary"]["n_iterations"]}
print(json.dumps(out))
