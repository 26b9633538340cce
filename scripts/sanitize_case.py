"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): shipped-image crop in both tie orders
(eager loop, CUDA-graph replay) and a 256^2 synthetic frame."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from detecting_cones_aoslo_b200 import configs
from detecting_cones_aoslo_b200.pipeline_with_delitions import find_blobs
from detecting_cones_aoslo_b200.synth import synth_aoslo, synth_config

d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "dataset.npz"))
img, gt = d["img"], d["gt"].astype(np.float64)
for tie in ("sklearn", "canonical"):
    cfg = configs.variant(configs.main_config(), tie_mode=tie, lambda_blob=1.0, dist_n_neighbours=3)
    res, _ = find_blobs(cfg, gt.copy(), img, None, False, False)
    print(tie, res["_summary"]["n_particles"], res["dice_coeff_2"])
im, g = synth_aoslo(256, 256, 5.8, seed=3)
for rep in range(2):
    cfg = synth_config(256, 256, 5.8, boundary=15, epoch_num=6, early_stop=False, tie_mode='canonical',
                       collect_stage_times=False)
    res, _ = find_blobs(cfg, g.copy(), im, None, False, False)
    print("synthetic", rep, res["_summary"]["n_particles"], res["dice_coeff_2"])
print("done")
