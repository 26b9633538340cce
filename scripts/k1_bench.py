"""K1 microbenchmark: the implementations / tilings of the fused blobness kernel on one synthetic frame.
CUDA events around 32 back-to-back launches over 8 output buffers (working set > L2), GPU parked while the host
enqueues; the deviation of every variant from the f64 band kernel is printed beside the time.

    python scripts/k1_bench.py [--size 2048] [--variants f64:32:2,f64:24:3,f32:auto,...]
variant = f64:<tile height>:<CTAs per SM> | f64:band | f32:<tile height>[:<occ>] | f32:auto | f32:band
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from detecting_cones_aoslo_b200.engine import Engine
from detecting_cones_aoslo_b200.synth import synth_aoslo

DEFAULTS = dict(k1_tma=1, pair_th=0, pair_occ=0, dtile_th=0, dtile_occ=0, dtile_split=1, k1_kernel=0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--variants", default="f64:band,f64:32:2,f64:24:2,f64:24:3,f64:16:2,f64:16:3,f64:16:4,f64:48:1,"
                                          "f32:auto,f32:band")
    args = ap.parse_args()
    b = 15
    img, _ = synth_aoslo(args.size, args.size, 5.8, 2022)
    pimg = np.pad(img, b, mode='symmetric')[: args.size + 2 * b - 1, : args.size + 2 * b - 1]   # odd width like 2077
    H, W = pimg.shape
    eng = Engine(H, W, particle_capacity=1024, gt_capacity=1024, device=0)
    d_img = eng.upload_image(pimg)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    eng.set_tuning(**dict(DEFAULTS, k1_kernel=2))
    ref_rb, ref_g = eng.blobness(d_img, 'custom', 'np_grad', (1.0,), torch.float64)
    ref_rb, ref_g = ref_rb.clone(), ref_g.clone()
    out = {"shape": [H, W], "results": {}}
    NB, REP = 8, 4
    for var in args.variants.split(","):
        f = var.split(":")
        knobs = dict(DEFAULTS)
        dt = torch.float64 if f[0] == "f64" else torch.float32
        bpp = 25.0 if f[0] == "f64" else 13.0
        if f[1] == "band":
            knobs["k1_kernel"] = 2
        elif f[0] == "f64":
            # f64:<th>:<occ>[:nosplit]
            knobs.update(k1_kernel=3, dtile_th=int(f[1]), dtile_occ=int(f[2]) if len(f) > 2 else 0,
                         dtile_split=0 if (len(f) > 3 and f[3] == "nosplit") else 1)
        elif f[1] != "auto":
            knobs.update(pair_th=int(f[1]), pair_occ=int(f[2]) if len(f) > 2 else 0)
        eng.set_tuning(**knobs)
        rbs = [eng.field(H, W, dt) for _ in range(NB)]
        gs = [eng.field(H, W, dt, 2) for _ in range(NB)]
        ms = []
        for it in range(2 + 3):
            flush.fill_(1)
            torch.cuda._sleep(6000000)
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(REP):
                for j in range(NB):
                    eng.blobness(d_img, 'custom', 'np_grad', (1.0,), dt, out=(rbs[j], gs[j]))
            e.record()
            torch.cuda.synchronize()
            if it >= 2:
                ms.append(a.elapsed_time(e) / (NB * REP))
        t = float(np.mean(ms))
        err_rb = float((rbs[0].double() - ref_rb).abs().max())
        err_g = float((gs[0].double() - ref_g).abs().max())
        out["results"][var] = {"us": t * 1e3, "GBps": bpp * H * W / t / 1e6, "max_err_Rb": err_rb, "max_err_grad": err_g}
        print("%-12s %8.2f us  %7.1f GB/s  err R_b %.2e grad %.2e" % (var, t * 1e3, bpp * H * W / t / 1e6, err_rb, err_g),
              flush=True)
        del rbs, gs
    print(json.dumps(out))


if __name__ == "__main__":
    main()
