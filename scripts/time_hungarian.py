"""Time of the device assignment solve (aoslo_hungarian, metric_algo='hangarian') on the shipped image's problem
(GT in the region x final particles of the config_zoo run), next to scipy on the host."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from detecting_cones_aoslo_b200 import configs
from detecting_cones_aoslo_b200.pipeline_with_delitions import find_blobs, prepare_inputs
from detecting_cones_aoslo_b200.utils import get_engine
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dataset.npz"))
img, gt = d["img"], d["gt"].astype(np.float64)
cfg = configs.zoo_config()
res, _ = find_blobs(dict(cfg), gt.copy(), img, None, False, False)
pos = res["_final_particles"]
pimg, pgt = prepare_inputs(cfg, gt.copy(), img)
eng = get_engine(*pimg.shape, n_gt=pgt.shape[0])
d_gt = eng.dev(pgt, torch.float64)
out = {"n_gt": int(d_gt.shape[0]), "n_particles": int(pos.shape[0])}
for _ in range(2):
    torch.cuda.synchronize(); t = time.perf_counter()
    s, m, v, _, steps = eng.hungarian(d_gt, pos)
    torch.cuda.synchronize(); out["gpu_s"] = time.perf_counter() - t
out.update({"sum": s, "mean": m, "var": v, "steps": steps, "us_per_step": out["gpu_s"] / steps * 1e6})
if "--scipy" in sys.argv:
    from scipy.optimize import linear_sum_assignment
    from scipy.spatial.distance import cdist
    dm = cdist(np.asarray(pgt, dtype=np.float64), pos.cpu().numpy())
    t = time.perf_counter(); r, c = linear_sum_assignment(dm); out["scipy_s"] = time.perf_counter() - t
    out["scipy_var"] = float(dm[r, c].var())
print(json.dumps(out))
