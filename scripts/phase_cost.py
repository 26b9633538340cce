"""Marginal cost of the phases of a run UNDER LANE CONCURRENCY (8 lanes on one GPU, the benchmark's throughput
setting): the same step as bench.py's `value`, timed for variants of the loop layout.  The differences say which
phase the throughput is paying for (kernel durations of a serialised launch list do not: small latency-bound
kernels overlap across lanes, grid-filling ones do not).

    python scripts/phase_cost.py [--lanes 8] [--steps 6]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", type=int, default=8)
    ap.add_argument("--steps", type=int, default=6)
    args = ap.parse_args()
    import torch
    import bench
    from concurrent.futures import ThreadPoolExecutor
    from detecting_cones_aoslo_b200.pipeline_with_delitions import make_config_struct
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lanes = [bench.Lane(j, dev) for j in range(args.lanes)]
    pool = ThreadPoolExecutor(max_workers=args.lanes)
    main_s = torch.cuda.current_stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step(fn):
        flush.fill_(1)
        torch.cuda._sleep(300000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main_s)
        for ln in lanes:
            ln.stream.wait_event(e0)
        for f in [pool.submit(fn, ln) for ln in lanes]:
            f.result()
        for ln in lanes:
            d = torch.cuda.Event()
            d.record(ln.stream)
            main_s.wait_event(d)
        e1.record(main_s)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def k1_only(ln):
        with torch.cuda.stream(ln.stream):
            ln.eng.blobness(ln.d_img, ln.cfg['blobness_formula'], ln.cfg['gradient_type'], (1.0,), torch.float64,
                            out=(ln.Rb, ln.grad))
            ln.stream.synchronize()

    variants = [
        ("full M=30 (resample every 4, metrics every 3)", dict()),
        ("M=2", dict(epoch_num=2)),
        ("M=30, metrics at iteration 0 only", dict(metric_measure_freq=1000)),
        ("M=30, no resample in the loop", dict(fix_particles_frequency=1000)),
        ("M=30, no resample, metrics at iteration 0 only", dict(fix_particles_frequency=1000, metric_measure_freq=1000)),
    ]
    out = {}
    for name, over in variants:
        for ln in lanes:
            cfg = dict(ln.cfg, **over)
            ln.cs = make_config_struct(cfg, ln.H, ln.W)
        for _ in range(3):
            step(lambda ln: ln.run_resident())
        ms = [step(lambda ln: ln.run_resident()) for _ in range(args.steps)]
        summ = lanes[0].last[0]
        out[name] = {"ms_per_run": float(np.mean(ms)) / args.lanes, "n_particles": summ.n_particles,
                     "n_seeds": summ.n_seeds}
        print("%-62s %.4f ms/run  (seeds %d, particles %d)" % (name, out[name]["ms_per_run"], summ.n_seeds,
                                                               summ.n_particles), flush=True)
    for _ in range(3):
        step(k1_only)
    ms = [step(k1_only) for _ in range(args.steps)]
    out["K1 alone"] = {"ms_per_run": float(np.mean(ms)) / args.lanes}
    print("%-62s %.4f ms/run" % ("K1 alone", out["K1 alone"]["ms_per_run"]))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
