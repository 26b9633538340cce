"""Readers for the fixtures written by oracle/gen_golden.py."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_dataset():
    d = np.load(os.path.join(GOLD, "dataset.npz"))
    return d["img"], d["gt"].astype(np.float64)


class GoldenTrace:
    """events: list of (kind, dict).  Sub-sampled fields carry `<key>_stats`,
    `<key>_shape`; `deleted` masks are unpacked to bool."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLD, "trace_%s.npz" % name))
        self.stride = int(z["stride"])
        kinds = [str(k) for k in z["kinds"]]
        per = {}
        for key in z.files:
            if not key.startswith("e") or key[4] != "_":
                continue
            per.setdefault(int(key[1:4]), {})[key[5:]] = z[key]
        self.events = []
        for i, kind in enumerate(kinds):
            p = per.get(i, {})
            if "deleted" in p:
                p["deleted"] = np.unpackbits(p["deleted"])[: int(p.pop("deleted_n"))].astype(bool)
            if kind == "acc":
                p["metrics"] = {t: {"TP": int(p["metrics_" + t][0]), "FP": int(p["metrics_" + t][1]),
                                    "FN": int(p["metrics_" + t][2]), "dice": float(p["metrics_" + t + "_f"][0]),
                                    "precision": float(p["metrics_" + t + "_f"][1]),
                                    "recall": float(p["metrics_" + t + "_f"][2])} for t in ("1", "2")}
            self.events.append((kind, p))
        self.result = {k[7:]: float(z[k]) for k in z.files if k.startswith("result_")}

    def of(self, kind):
        return [p for k, p in self.events if k == kind]
