"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by EXECUTING THE REFERENCE.

Run in the build container (needs /root/reference):

    python -m oracle.gen_golden            # all fixtures
    python -m oracle.gen_golden main zoo   # a subset

Every fixture is the trace of one `find_blobs` run of the unmodified reference
(oracle/reference_harness.py: reference modules + matplotlib/imageio/skimage
stubs), i.e. outputs of the reference itself -- this is what pins the oracle.
The shipped image + CSV are stored as arrays in tests/golden/dataset.npz so the
GPU box (which has no /root/reference) can replay the same inputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import reference_harness as rh                      # noqa: E402
from detecting_cones_aoslo_b200 import configs                   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# name -> (config factory, as_shipped_energy (False = closed form, identical outputs), field stride)
RUNS = {
    "main": (configs.main_config, True, 1),
    "main_sobel": (lambda: configs.variant(configs.main_config(), gradient_type="sobel", lambda_blob=1.), True, 1),
    "main_contin": (lambda: configs.variant(configs.main_config(), step_mode="contin", lambda_blob=1.), True, 1),
    "main_exp": (lambda: configs.variant(configs.main_config(), dist_energy_type="exp"), True, 1),
    "main_hangarian": (lambda: configs.variant(configs.main_config(), metric_algo="hangarian"), True, 1),
    "main_k3": (lambda: configs.variant(configs.main_config(), dist_n_neighbours=3, lambda_blob=1.,
                                        lambda_dist=0.1), True, 1),
    "zoo": (configs.zoo_config, True, 4),
    "full_b10": (configs.full_b10_config, False, 4),
    "sweep_like": (configs.sweep_like_config, False, 4),
}


def _compact(a):
    a = np.asarray(a)
    if a.dtype.kind == "f" and a.size and np.isfinite(a).all() and (a == np.round(a)).all() \
            and np.abs(a).max() < 32000:
        return a.astype(np.int16)
    if a.dtype == np.int64 and a.size and np.abs(a).max() < 32000:
        return a.astype(np.int16)
    return a


def pack_trace(trace, result, stride):
    out = {}
    kinds = []
    for i, (kind, payload) in enumerate(trace.events):
        kinds.append(kind)
        for key, val in payload.items():
            name = "e%03d_%s" % (i, key)
            if key == "metrics":
                for t, m in val.items():
                    out[name + "_" + t] = np.array([m["TP"], m["FP"], m["FN"]], dtype=np.int64)
                    out[name + "_" + t + "_f"] = np.array([m["dice"], m["precision"], m["recall"]],
                                                          dtype=np.float64)
                continue
            val = np.asarray(val)
            if key in ("R_b", "grad", "field") and stride > 1:
                out[name + "_stats"] = np.array([val.min(), val.max(), val.mean()], dtype=np.float64)
                out[name + "_shape"] = np.array(val.shape[:2], dtype=np.int64)
                val = val[::stride, ::stride]
            if key == "deleted":
                out[name + "_n"] = np.array(val.size, dtype=np.int64)
                val = np.packbits(val.astype(np.uint8))
            if key == "gt" and i > 0 and any(k.endswith("_gt") for k in out):
                continue                                   # GT is constant during a run; keep one copy
            out[name] = _compact(val)
    out["kinds"] = np.array(kinds)
    out["stride"] = np.array(stride)
    for k in ("best_lsa_mean", "best_blobEn_mean", "dice_coeff_1", "dice_coeff_2"):
        out["result_" + k] = np.array(float(result[k]), dtype=np.float64)
    return out


def main(names):
    os.makedirs(GOLD, exist_ok=True)
    img, gt = rh.load_dataset()
    np.savez_compressed(os.path.join(GOLD, "dataset.npz"), img=img, gt=gt)
    for name in names:
        factory, as_shipped, stride = RUNS[name]
        cfg = factory()
        tr = rh.Trace()
        result, _, tr = rh.run_find_blobs(cfg, img=img, gt=gt, trace=tr, closed_form=not as_shipped)
        packed = pack_trace(tr, result, stride)
        path = os.path.join(GOLD, "trace_%s.npz" % name)
        np.savez_compressed(path, **packed)
        n_del = [p["particles_out"].shape[0] for p in tr.of("delete")]
        print("%-14s events=%d  deletes->%s  dice2=%r  %.0f KB" % (
            name, len(tr.events), n_del, float(result["dice_coeff_2"]), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main(sys.argv[1:] or list(RUNS))
