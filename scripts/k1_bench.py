"""K1 microbenchmark: every implementation of the fused blobness kernel on one synthetic frame,
CUDA events, L2 flushed between launches, plus the deviation of each f32 variant from the f64 kernel.

    python scripts/k1_bench.py [--size 2048] [--iters 20]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--flush", default="write", choices=["write", "write+read", "none"],
                    help="L2 state before each timed launch: 256 MiB write (dirty lines), the same followed by a "
                         "256 MiB read sweep (cold and clean), or nothing (warm)")
    ap.add_argument("--dbg", default="0")
    ap.add_argument("--variants", default="band,pair:16,pair:20,pair:24,pair:28,pair:32,pair:36,pair:24:2,pair:auto")
    args = ap.parse_args()
    os.environ["AOSLO_K1_DBG"] = args.dbg
    b = 15
    img, _ = synth_aoslo(args.size, args.size, 5.8, 2022)
    pimg = np.pad(img, b, mode='symmetric')[: args.size + 2 * b - 1, : args.size + 2 * b - 1]   # odd width like 2077
    H, W = pimg.shape
    eng = Engine(H, W, particle_capacity=1024, gt_capacity=1024, device=0)
    d_img = eng.upload_image(pimg)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
    flush2 = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device='cuda')
    Rb64 = eng.field(H, W, torch.float64)
    g64 = eng.field(H, W, torch.float64, 2)
    os.environ["AOSLO_BLOBNESS_KERNEL"] = "band"
    eng.blobness(d_img, 'custom', 'np_grad', (1.0,), torch.float64, out=(Rb64, g64))
    torch.cuda.synchronize()
    out = {"shape": [H, W], "flush": args.flush, "results": {}}
    for var in args.variants.split(","):
        os.environ.pop("AOSLO_BLOBNESS_KERNEL", None)
        os.environ.pop("AOSLO_PAIR_TH", None)
        os.environ.pop("AOSLO_PAIR_OCC", None)
        os.environ.pop("AOSLO_DTILE_TH", None)
        f64 = var.startswith("d:")
        dt = torch.float64 if f64 else torch.float32
        bpp = 25.0 if f64 else 13.0
        if var == "d:band":
            os.environ["AOSLO_BLOBNESS_KERNEL"] = "band"
        elif f64:
            os.environ["AOSLO_BLOBNESS_KERNEL"] = "dtile"
            os.environ["AOSLO_DTILE_TH"] = var.split(":")[1]
        elif var == "band":
            os.environ["AOSLO_BLOBNESS_KERNEL"] = "band"
        elif var.startswith("pair:") and var != "pair:auto":
            os.environ["AOSLO_PAIR_TH"] = var.split(":")[1]
            if len(var.split(":")) > 2:
                os.environ["AOSLO_PAIR_OCC"] = var.split(":")[2]
        Rb = eng.field(H, W, dt)
        g = eng.field(H, W, dt, 2)
        Rb.fill_(float('nan'))
        g.fill_(float('nan'))
        ms = []
        for i in range(3 + args.iters):
            if args.flush != "none":
                flush.fill_(1)
            if args.flush == "write+read":
                sink = flush2.sum()
            torch.cuda._sleep(200000)
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.blobness(d_img, 'custom', 'np_grad', (1.0,), dt, out=(Rb, g))
            e.record()
            torch.cuda.synchronize()
            if i >= 3:
                ms.append(a.elapsed_time(e))
        # streamed: NB distinct frames (inputs + outputs 8 x 56 MB > the 126 MB L2) back to back between ONE event
        # pair, the GPU parked behind a sleep kernel while the host enqueues, so that neither launch latency nor
        # the ~1 us event resolution enters the per-launch figure
        NB, REP = 8, 4
        if not hasattr(main, "_bufs"):
            main._bufs = {}
        if dt not in main._bufs:
            main._bufs.clear()
            main._bufs[dt] = ([eng.upload_image(pimg) for _ in range(NB)],
                              [eng.field(H, W, dt) for _ in range(NB)],
                              [eng.field(H, W, dt, 2) for _ in range(NB)])
        imgs, rbs, gs = main._bufs[dt]
        stream_us = []
        for it in range(3):
            flush.fill_(1)
            torch.cuda._sleep(6000000)
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for rep in range(REP):
                for f in range(NB):
                    eng.blobness(imgs[f], 'custom', 'np_grad', (1.0,), dt, out=(rbs[f], gs[f]))
            e.record()
            torch.cuda.synchronize()
            stream_us.append(a.elapsed_time(e) * 1e3 / (NB * REP))
        sus = float(np.min(stream_us))
        # the same stream of frames issued round-robin on two CUDA streams: the launch ramp / tail of one
        # launch overlaps the body of the next
        s2_us = []
        if not hasattr(main, "_streams"):
            main._streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        cur = torch.cuda.current_stream()
        for it in range(3):
            flush.fill_(1)
            torch.cuda._sleep(6000000)
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for st_ in main._streams:
                st_.wait_event(a)
            for rep in range(REP):
                for f in range(NB):
                    with torch.cuda.stream(main._streams[f & 1]):
                        eng.blobness(imgs[f], 'custom', 'np_grad', (1.0,), dt, out=(rbs[f], gs[f]))
            for st_ in main._streams:
                d = torch.cuda.Event()
                d.record(st_)
                cur.wait_event(d)
            e.record()
            torch.cuda.synchronize()
            s2_us.append(a.elapsed_time(e) * 1e3 / (NB * REP))
        st = eng.status(clear=True)
        err_rb = float((Rb.double() - Rb64).abs().max() / Rb64.abs().max())
        err_g = float((g.double() - g64).abs().max() / g64.abs().max())
        if f64:          # against the band kernel: must be bit identical
            err_rb = float((Rb != Rb64).sum())
            err_g = float((g != g64).sum())
        us = float(np.median(ms)) * 1e3
        out["results"][var] = {"us_median": us, "us_min": float(np.min(ms)) * 1e3, "gbs": bpp * H * W / us / 1e3, "bytes_per_px": bpp,
                               "frac_of_6459.6": bpp * H * W / us / 1e3 / 6459.6, "rel_err_Rb": err_rb,
                               "rel_err_grad": err_g, "status": st, "stream_us": sus, "stream_all": stream_us, "two_streams_us": s2_us,
                               "stream_frac_of_6459.6": bpp * H * W / sus / 1e3 / 6459.6,
                               "nan": bool(torch.isnan(Rb).any() or torch.isnan(g).any())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
