#!/usr/bin/env python
"""bench.py -- detection runs/s of the full `find_blobs` pipeline on a synthetic 2048^2 AOSLO frame.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A step = `--batch` whole detection runs (blobness field + gradient, seeding, thinning, resampling,
M move/delete iterations, metrics), each on its own synthetic frame, run concurrently on one GPU
(one CUDA stream + scratch context per frame -- frames / sweep trials are independent).
Per-GPU work is fixed as N grows (weak scaling); the only collective is one NCCL all-gather of the
per-run metric records at the end of the timed region.  The single-run latency is reported too.

value     runs/s with the frame and GT already resident in HBM (timed: aoslo_blobness + aoslo_run)
e2e       runs/s through the reference-facing entry point find_blobs(cur_config, GT_data, img, ...)
          with HOST (pinned) NumPy buffers: host->device copies of the raw frame and GT, crop/pad on the device
          and the device->host read of the metric records are inside the timed region
roofline  the fused blobness kernel (K1) in the f64 form the timed step launches (25 B/px), CUDA events around the
          kernel inside the timed single-lane step, against MEASURED_PEAKS.json; roofline_f32 is the f32 form
          north_star states the field tolerance for (13 B/px), timed around 32 back-to-back launches over 8
          distinct frames (> L2)
faithful  the drop-in's default mode (sklearn KD-tree tie order, early stop on), end to end
cpu_baseline / --impl reference
          the CPU oracle (port of the reference's NumPy/SciPy/scikit-learn pipeline; the reference
          is Python and cannot travel to the GPU box) on the SAME full 2048^2 frame, configuration, tie order
          and iteration count: 1 process (cpu_baseline) resp. one process per core, <= 16 (--impl reference)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BOUNDARY = 15
TIE_MODE = 'canonical'  # both arms: (distance, index) neighbour order; the as-is reference order ('sklearn') is
EARLY_STOP = False      # reported beside it as `faithful`; all M iterations run in both arms
FALLBACK_HBM_GBS = 6650.0

# BASELINE.json configs.  `default` is the one the metric is quoted on ("detection runs/s, full pipeline, 2048^2
# img"); cfg2 / cfg3 are configs[1] / [2]; `batch` / `sweep` (configs[3] / [4]) run through sweep.UnitRunner.
WORKLOADS = {
    'default': dict(size=2048, pitch=5.8, epochs=30, scales=(1.0,), lanes=8,
                    metric="detection runs/s (full pipeline, 2048^2 image)"),
    'cfg2': dict(size=1024, pitch=11.0, epochs=100, scales=(1.0, 1.5, 2.0, 2.5, 3.0), lanes=8,
                 metric="detection runs/s (cfg2: 1024^2, 5 blobness scales, M=100)"),
    'cfg3': dict(size=4096, pitch=11.4, epochs=30, scales=(1.0,), lanes=4,
                 metric="detection runs/s (cfg3: 4096^2 montage)"),
    'batch': dict(size=512, pitch=5.8, epochs=30, scales=(1.0,), lanes=8, units=256,
                  metric="frames/s (cfg4: batch of 256 synthetic 512^2 frames, sharded by image)"),
    'sweep': dict(size=2048, pitch=5.8, epochs=30, scales=(1.0,), lanes=8, units=1024,
                  metric="trials/s (cfg5: 1024 sweep trials of the sweep_config_new.yaml space on one 2048^2 image)"),
}
WL = WORKLOADS['default']
SIZE, PITCH, EPOCHS = WL['size'], WL['pitch'], WL['epochs']


def select_workload(name):
    global WL, SIZE, PITCH, EPOCHS
    WL = WORKLOADS[name]
    SIZE, PITCH, EPOCHS = WL['size'], WL['pitch'], WL['epochs']


def workload_name():
    sc = "" if len(WL['scales']) == 1 else ", blobness scales %s" % (list(WL['scales']),)
    return ("synthetic %dx%d AOSLO-like cone lattice (pitch %.1f px, seed 2022+rank), config_zoo, boundary %d, "
            "k=3, M=%d iterations (early stop off), f64 fields, canonical tie order%s"
            % (SIZE, SIZE, PITCH, BOUNDARY, EPOCHS, sc))


def make_workload(rank, size=None):
    from detecting_cones_aoslo_b200.synth import synth_aoslo, synth_config
    size = size or SIZE
    img, gt = synth_aoslo(size, size, PITCH, seed=2022 + rank)
    cfg = synth_config(size, size, PITCH, boundary=BOUNDARY, epoch_num=EPOCHS, early_stop=EARLY_STOP,
                       tie_mode=TIE_MODE, field_dtype='f64', collect_stage_times=False)
    if len(WL['scales']) > 1:
        cfg['blobness_scales'] = tuple(WL['scales'])
    return img, gt, cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin=None):
        """t_begin: start of the timed region (perf_counter).  The sampler is started before the warm-up (nvidia-smi
        needs ~0.5 s to come up and the timed region of a multi-GPU run is a few tens of ms); samples taken inside the
        timed region are used when there are any, otherwise the ones of warm-up + timed region (same load)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [ln for t, ln in self.lines if t_begin is not None and t >= t_begin - 0.05]
        window = "timed region" if inside else "warm-up + timed region"
        for ln in (inside or [ln for _, ln in self.lines]):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def k1_dram_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per K1 launch from the committed ncu --set full capture
    (profiles/r02_k1_traffic.json, else r01; written by scripts/summarise_profiles.py); None if absent."""
    for name in ("r02_k1_traffic.json", "r01_k1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)["dram_bytes_per_launch"][key])
        except Exception:
            continue
    return None


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------- CPU arm
def cpu_full_run(seed_index=0, tie_mode=TIE_MODE, early_stop=EARLY_STOP):
    """One oracle run (NumPy / SciPy / scikit-learn port of the reference's find_blobs, closed-form pit energy) on
    the SAME 2048^2 frame, configuration, tie order and iteration count as the GPU arm."""
    from oracle import aoslo_oracle as orc
    img, gt, cfg = make_workload(seed_index)
    ocfg = {k: v for k, v in cfg.items() if k not in ('early_stop', 'tie_mode', 'field_dtype', 'collect_stage_times',
                                                      'blobness_scales')}
    t = time.perf_counter()
    res, ts = orc.find_blobs(ocfg, gt, img, None, False, False, tie_mode=tie_mode, early_stop=early_stop,
                             scales=tuple(WL['scales']))
    return time.perf_counter() - t, res, ts


def _cpu_worker(i):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    return cpu_full_run(0)[0]


def cpu_side_rows():
    """The two extra CPU rows SURVEY 8d asks for: the blobness stage alone (MPix/s next to `blobness_mpix_s`) and the
    reference AS SHIPPED, whose pit energy builds two scipy.stats frozen distributions per call (utils.py:82-87): the
    per-call cost is measured on 2 000 calls and multiplied by the number of calls the run makes."""
    from oracle import aoslo_oracle as orc
    from detecting_cones_aoslo_b200.pipeline_with_delitions import prepare_inputs
    import scipy.stats
    img, gt, cfg = make_workload(0)
    pimg, _ = prepare_inputs(cfg, gt.copy(), img)
    best = 1e9
    for _ in range(2):
        t = time.perf_counter()
        rb = orc.calc_blobness(pimg, 'custom')
        orc.calc_grad_field(rb, 'np_grad')
        best = min(best, time.perf_counter() - t)
    d = np.linspace(0.5, 12.0, 2000)
    t = time.perf_counter()
    for x in d:                                          # energy_pit_func2 as the reference ships it
        ef = scipy.stats.expon(cfg['lambda_dist_func'])
        nf = scipy.stats.norm(cfg['mu_dist_func'], cfg['sigma_dist_func'])
        _ = ef.pdf(x) - nf.pdf(x) + 0.48 if x <= cfg['mu_dist_func'] else ef.pdf(x) + nf.pdf(x) - 0.48
    per_call = (time.perf_counter() - t) / d.size
    return {"blobness_mpix_s": pimg.size / best / 1e6, "blobness_seconds": best,
            "as_shipped_energy_call_us": per_call * 1e6}


def cpu_baseline(processes=1, side_rows=False):
    """runs/s of the CPU oracle on the full frame: `processes` concurrent single-threaded runs."""
    if processes <= 1:
        dt, res, ts = cpu_full_run(0)
        runs_per_s = 1.0 / dt
        n_calls = res.get('n_energy_calls')
    else:
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        t = time.perf_counter()
        with ctx.Pool(processes) as pool:
            pool.map(_cpu_worker, range(processes))
        dt = time.perf_counter() - t
        runs_per_s = processes / dt
        n_calls = None
    out = {"value": runs_per_s, "unit": "runs/s", "cores": processes, "kind": "port",
           "sample": "CPU oracle (NumPy/SciPy/scikit-learn port of the reference's find_blobs, closed-form pit "
                     "energy, %s tie order) on the full %dx%d frame of the GPU arm, same config, M=%d (early stop "
                     "off): 1 whole run per process, %d concurrent process(es), %.1f s wall"
                     % (TIE_MODE, SIZE, SIZE, EPOCHS, processes, dt),
           "seconds_per_sample": dt}
    if side_rows:
        rows = cpu_side_rows()
        out["blobness_mpix_s"] = rows["blobness_mpix_s"]
        if n_calls:
            shipped_s = dt + n_calls * rows["as_shipped_energy_call_us"] * 1e-6
            out["as_shipped"] = {"value": 1.0 / shipped_s, "unit": "runs/s",
                                 "how": "closed-form run time + %d energy_pit_func2 calls x %.0f us (two scipy.stats "
                                        "frozen distributions per call, utils.py:82-87; measured on 2000 calls)"
                                        % (n_calls, rows["as_shipped_energy_call_us"])}
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    procs = max(1, min(os.cpu_count() or 1, 16))
    t0 = time.perf_counter()
    vals = []
    steps = max(1, min(args.steps, 2))          # each step = `procs` whole 2048^2 CPU runs (~2 min); keep the arm to minutes
    base = None
    for _ in range(steps):
        base = cpu_baseline(processes=procs)
        vals.append(base["value"])
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": WL['metric'], "value": v,
        "unit": "runs/s", "n_gpus": args.gpus, "steps": steps, "warmup": 0,
        "ms_per_step": base["seconds_per_sample"] * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(), "l2": "n/a (CPU arm)"},
        "cpu_baseline": dict(base, value=v),
        "e2e": {"value": v, "unit": "runs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------- GPU arm
class Lane:
    """One concurrent detection lane: its own synthetic frame, context (scratch) and CUDA stream."""

    def __init__(self, frame_seed_index, dev):
        import torch
        from detecting_cones_aoslo_b200.engine import Engine
        from detecting_cones_aoslo_b200.pipeline_with_delitions import prepare_inputs, make_config_struct
        self.torch = torch
        self.img, self.gt, self.cfg = make_workload(frame_seed_index)
        # the e2e leg hands find_blobs HOST arrays that live in pinned memory (the frame, and a scratch copy of
        # the GT that is refreshed before every call because find_blobs shifts its GT argument in place)
        self._img_pin = torch.from_numpy(self.img).pin_memory()
        self.img = self._img_pin.numpy()
        self._gt_pin = torch.from_numpy(self.gt.copy()).pin_memory()
        self.gt_scratch = self._gt_pin.numpy()
        pimg, pgt = prepare_inputs(self.cfg, self.gt.copy(), self.img)
        self.H, self.W = pimg.shape
        self.h2d = self.img.nbytes + self.gt.nbytes       # what find_blobs copies host->device per call
        self.eng = Engine(self.H, self.W, particle_capacity=self.H * self.W, gt_capacity=max(65536, self.gt.shape[0]),
                          device=dev.index)
        self.stream = torch.cuda.Stream(device=dev)
        self.pimg = pimg
        self.d_img = self.eng.upload_image(pimg)
        self.d_gt = self.eng.dev(pgt, torch.float64)
        self.cs = make_config_struct(self.cfg, self.H, self.W)
        self.Rb = self.eng.field(self.H, self.W, torch.float64)
        self.grad = self.eng.field(self.H, self.W, torch.float64, 2)
        self.n_gt = int(pgt.shape[0])
        self.k1_events = []
        self.last = None
        self.collect_times = False       # stage timings off: aoslo_run replays its cached CUDA graph
        self._gt_steps = []

    def run_resident(self, time_k1=False):
        """frame + GT resident in HBM: aoslo_blobness + aoslo_run (the whole loop)."""
        from detecting_cones_aoslo_b200 import _lib
        torch = self.torch
        with torch.cuda.stream(self.stream):
            self.eng.clear_status()                      # stream ordered; read back with the run's summary
            if time_k1:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            self.eng.blobness(self.d_img, self.cfg['blobness_formula'], self.cfg['gradient_type'],
                              tuple(WL['scales']), torch.float64, out=(self.Rb, self.grad))
            if time_k1:
                e1.record()
                self.k1_events.append((e0, e1))
            summ, recs, pos, st, ms = self.eng.run(self.cs, self.Rb, self.grad, self.d_gt, collect_times=self.collect_times)
        if summ.status:
            _lib.raise_device_status(summ.status)
        self.last = (summ, recs, ms)
        return summ

    def run_e2e(self, gt_host=None, **over):
        """the reference-facing call with HOST buffers (H2D of the raw frame + GT, device crop/pad, run, D2H of the
        records)."""
        from detecting_cones_aoslo_b200.pipeline_with_delitions import find_blobs
        gt_host = self.gt_scratch if gt_host is None else gt_host
        with self.torch.cuda.stream(self.stream):
            res, _ = find_blobs(dict(self.cfg, **over), gt_host, self.img, None, False, False, engine=self.eng)
        return res

    def prep_e2e_steps(self, k):
        """host inputs of k consecutive end-to-end steps without a barrier in between: k pinned GT copies (find_blobs
        shifts its GT argument in place), made before the timed region"""
        k = min(k, self.MAX_GT_COPIES)
        while len(self._gt_steps) < k:
            self._gt_steps.append(self.torch.from_numpy(self.gt.copy()).pin_memory())
        for t in self._gt_steps[:k]:
            np.copyto(t.numpy(), self.gt)
        return [t.numpy() for t in self._gt_steps[:k]]

    MAX_GT_COPIES = 32       # pinned memory bound: beyond 32 steps a copy is restored inside the timed loop (counted)

    def e2e_step(self, copies, i):
        """step i of the end-to-end region on this lane"""
        g = copies[i % len(copies)]
        if i >= len(copies):
            np.copyto(g, self.gt)
        return self.run_e2e(g)

    def prep_e2e(self):
        """this step's host input: find_blobs shifts the caller's GT array in place (reference :26), so the bench
        hands it a fresh copy every step -- made before the timed region, like the L2 flush"""
        np.copyto(self.gt_scratch, self.gt)


def pin_rank_to_core_slice(local, world):
    """Every rank keeps its lane threads on its own slice of the host cores (set before the threads exist): 8 ranks x
    8 lane threads on 32 cores otherwise migrate across all of them -- measured on the 8-GPU box: 6 689 -> 6 796
    runs/s resident, 5 229 -> 6 014 end to end.  AOSLO_BENCH_AFFINITY=0 turns it off."""
    if world > 1 and os.environ.get("AOSLO_BENCH_AFFINITY", "1") != "0" and hasattr(os, "sched_setaffinity"):
        avail = sorted(os.sched_getaffinity(0))
        per = max(1, len(avail) // world)
        os.sched_setaffinity(0, avail[local * per:(local + 1) * per] or avail)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from concurrent.futures import ThreadPoolExecutor
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    from detecting_cones_aoslo_b200 import _lib
    from detecting_cones_aoslo_b200.utils import metrics_from_record
    cores = os.cpu_count() or 1
    # lanes per GPU: every lane has a host thread that enqueues one run (a CUDA graph + ~90 eager launches)
    # and waits for the device.  8 lanes saturate one B200 (1/2/4/8 lanes: 295/450/600/690 runs/s) and were
    # also the best setting on the 8-GPU / 32-core box (4388 runs/s vs 3826 with 4 lanes per GPU).
    # lanes per GPU: 8 saturate most of one B200 (4 / 8 / 12 / 16 / 24 lanes: 680 / 845 / 863 / 878 / 904 runs/s, e2e
    # 550 / 696 / 743 / 754 / 804 -- more lanes mainly hide the host side of the end-to-end path).  The default stays 8
    # at every N so that the per-GPU configuration of the weak-scaling run is the same (8 lanes per GPU were also the
    # best setting on the 8-GPU / 32-core box).
    B = args.batch if args.batch > 0 else max(1, min(WL['lanes'], cores))
    # host wait policy: spin unless the lane threads of all ranks oversubscribe the cores, then yield
    # (measured on the 8-GPU / 32-core box before the CUDA-graph loop: 2938 runs/s with yield vs 2639 with spin)
    policy = os.environ.get("AOSLO_BENCH_WAIT_POLICY", "2" if world * B * 2 > cores else "")
    if policy in ("1", "2", "3"):
        st = _lib.lib.aoslo_set_wait_policy(local, int(policy))       # before the context is created
        if st != 0:
            sys.stderr.write("aoslo_set_wait_policy failed: %s\n" % _lib.lib.aoslo_last_cuda_error().decode())
    pin_rank_to_core_slice(local, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    lanes = [Lane(rank * B + j, dev) for j in range(B)]
    if os.environ.get("AOSLO_BENCH_BLOCKING_SYNC", "0") == "1":
        for ln in lanes:                 # measured slower than spinning on the 32-core box (wake-up latency)
            ln.eng.set_blocking_sync(True)
    H, W = lanes[0].H, lanes[0].W
    pool = ThreadPoolExecutor(max_workers=B)
    main = torch.cuda.current_stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def evict_l2():
        """Evict L2 (256 MiB write) and park the GPU ~150 us so the host runs ahead of the device:
        the events then bracket device work, not Python launch latency.  Untimed."""
        flush.fill_(1)
        torch.cuda._sleep(300000)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_step(fn, active, prep=None):
        """one step = every active lane once, concurrently; returns device ms (CUDA events on `main`,
        lane streams forked from / joined to it).  `prep` builds the step's host inputs, untimed."""
        if prep is not None:
            for ln in active:
                prep(ln)
        evict_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for ln in active:
            ln.stream.wait_event(e0)
        if len(active) == 1:
            fn(active[0])
        else:
            for f in [pool.submit(fn, ln) for ln in active]:
                f.result()
        for ln in active:
            d = torch.cuda.Event()
            d.record(ln.stream)
            main.wait_event(d)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def timed_steps(fn, active, k):
        """k steps back to back inside ONE bracket (the contract's form): every lane runs its k frames one after the
        other, no barrier and no L2 flush between steps -- the lanes' working sets (8 x ~250 MB written per run)
        are far larger than the 126 MB L2, so every run reads its frame from HBM.  fn(lane, step_index).  Returns
        device ms of the whole region (CUDA events on `main`, lane streams forked from / joined to it)."""
        evict_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for ln in active:
            ln.stream.wait_event(e0)

        def loop(ln):
            for i in range(k):
                fn(ln, i)
        for f in [pool.submit(loop, ln) for ln in active]:
            f.result()
        for ln in active:
            d = torch.cuda.Event()
            d.record(ln.stream)
            main.wait_event(d)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # ---- warm-up (all lanes), then the throughput pass: K steps of B concurrent runs
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        timed_step(lambda ln: ln.run_resident(), lanes)
    barrier()
    t_timed = time.perf_counter()
    launches0 = _lib.lib.aoslo_launch_count()
    region_ms = timed_steps(lambda ln, i: ln.run_resident(), lanes, args.steps)
    launches_region = _lib.lib.aoslo_launch_count() - launches0
    # the same K steps one by one, a barrier + L2 flush between them (how rounds 1 / 2 timed the step): `stepwise`
    step_ms = [timed_step(lambda ln: ln.run_resident(), lanes) for _ in range(args.steps)]
    # the sweep's only exchange: all-gather of the per-run metric records (tiny, latency only)
    rows = []
    for ln in lanes:
        summ, recs, _ms = ln.last
        pm, lsa_sum, lsa_mean, lsa_var, blob_mean = metrics_from_record(recs[-1])
        rows.append([rank, summ.n_particles, summ.n_iterations, float(pm['1']['dice']), float(pm['2']['dice']),
                     lsa_mean, blob_mean, summ.n_seeds])
    rec_t = torch.tensor(rows, dtype=torch.float64, device=dev)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    if world > 1:
        gathered = [torch.zeros_like(rec_t) for _ in range(world)]
        dist.all_gather(gathered, rec_t)
    else:
        gathered = [rec_t]
    g1.record()
    barrier()
    launches = launches_region
    clocks = sampler.stop(t_timed)
    tt = torch.tensor([region_ms + g0.elapsed_time(g1), float(sum(step_ms)) + g0.elapsed_time(g1)],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms, stepwise_ms = float(tt[0].item()), float(tt[1].item())
    value = world * args.steps * B / (total_ms / 1e3)
    value_stepwise = world * args.steps * B / (stepwise_ms / 1e3)

    # ---- latency pass: one lane alone (single stream, spinning waits), K1 timed inside the step
    lane0 = lanes[0]
    lane0.eng.set_blocking_sync(False)
    for _ in range(2):
        timed_step(lambda ln: ln.run_resident(), [lane0])
    lane0.k1_events.clear()
    lat_ms = [timed_step(lambda ln: ln.run_resident(time_k1=True), [lane0]) for _ in range(args.steps)]
    k1_ms = float(np.mean([a.elapsed_time(b) for a, b in lane0.k1_events]))
    latency_ms = float(np.mean(lat_ms))
    lane0.collect_times = True                      # one eager run for the reference's five stage buckets
    lane0.run_resident()
    summ, recs, stage_ms = lane0.last
    lane0.collect_times = False

    # ---- end to end through the reference-facing entry point, HOST buffers in, records out
    for _ in range(min(2, args.warmup)):
        timed_step(lambda ln: ln.run_e2e(), lanes, Lane.prep_e2e)
    barrier()
    gt_steps = {id(ln): ln.prep_e2e_steps(args.steps) for ln in lanes}       # untimed, like the L2 flush
    e2e_region_ms = timed_steps(lambda ln, i: ln.e2e_step(gt_steps[id(ln)], i), lanes, args.steps)
    barrier()
    e2e_ms = [timed_step(lambda ln: ln.run_e2e(), lanes, Lane.prep_e2e) for _ in range(args.steps)]
    te = torch.tensor([e2e_region_ms, float(sum(e2e_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps * B / (float(te[0].item()) / 1e3)
    e2e_stepwise = world * args.steps * B / (float(te[1].item()) / 1e3)
    e2e_lat = [timed_step(lambda ln: ln.run_e2e(), [lane0], Lane.prep_e2e) for _ in range(args.steps)]
    d2h = summ.n_records * 88 + 32 + 4

    # ---- blobness MPix/s (second half of BASELINE.json's metric): K1 alone, f32 fields (north_star's field
    # tolerance is stated in fp32; 13 B/px).  Two figures:
    #   single   one launch between two events, L2 flushed before it (event resolution ~1 us, launch ramp inside);
    #   stream   8 distinct frames (inputs + outputs 8 x 56 MB > the 126 MB L2, so every launch reads its bytes from
    #            HBM and its stores push older frames out) x 4 rounds back to back between ONE event pair, the GPU
    #            parked behind a sleep kernel while the host enqueues: the kernel's average launch duration
    eng = lane0.eng
    Rb32 = eng.field(H, W, torch.float32)
    g32 = eng.field(H, W, torch.float32, 2)
    f32_ms = []
    for i in range(3 + 10):
        evict_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        eng.blobness(lane0.d_img, 'custom', 'np_grad', (1.0,), torch.float32, out=(Rb32, g32))
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            f32_ms.append(a.elapsed_time(b))
    k1_f32_ms = float(np.mean(f32_ms))
    NB, REP = 8, 4
    s_img = [eng.upload_image(lanes[j % len(lanes)].pimg) for j in range(NB)]
    s_rb = [eng.field(H, W, torch.float32) for _ in range(NB)]
    s_g = [eng.field(H, W, torch.float32, 2) for _ in range(NB)]
    stream_ms = []
    for i in range(2 + 3):
        evict_l2()
        torch.cuda._sleep(6000000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(REP):
            for j in range(NB):
                eng.blobness(s_img[j], 'custom', 'np_grad', (1.0,), torch.float32, out=(s_rb[j], s_g[j]))
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            stream_ms.append(a.elapsed_time(b) / (NB * REP))
    k1_f32_stream_ms = float(np.mean(stream_ms))
    # the same 32 launches alternating on two CUDA streams (what concurrent lanes do): the ramp / tail of one launch
    # overlaps the body of the next
    two = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    two_ms = []
    for i in range(2 + 3):
        evict_l2()
        torch.cuda._sleep(6000000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for st_ in two:
            st_.wait_event(a)
        for _ in range(REP):
            for j in range(NB):
                with torch.cuda.stream(two[j & 1]):
                    eng.blobness(s_img[j], 'custom', 'np_grad', (1.0,), torch.float32, out=(s_rb[j], s_g[j]))
        for st_ in two:
            d = torch.cuda.Event()
            d.record(st_)
            main.wait_event(d)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            two_ms.append(a.elapsed_time(b) / (NB * REP))
    k1_f32_two_ms = float(np.mean(two_ms))
    # the f64 form (the kernel the timed step runs), same protocol: 32 back-to-back launches over 8 distinct frames
    d_rb = [eng.field(H, W, torch.float64) for _ in range(NB)]
    d_g = [eng.field(H, W, torch.float64, 2) for _ in range(NB)]
    f64_ms = []
    for i in range(2 + 3):
        evict_l2()
        torch.cuda._sleep(6000000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(REP):
            for j in range(NB):
                eng.blobness(s_img[j], 'custom', 'np_grad', tuple(WL['scales']), torch.float64, out=(d_rb[j], d_g[j]))
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            f64_ms.append(a.elapsed_time(b) / (NB * REP))
    k1_f64_stream_ms = float(np.mean(f64_ms))
    del s_img, s_rb, s_g, d_rb, d_g

    # ---- the as-is reference semantics: sklearn KD-tree tie order (host-built trees, one round trip per kNN) and
    # the early stop on -- the drop-in's default mode, the one the golden traces of the executed reference pin
    faithful = None
    if not args.no_faithful:
        fo = dict(tie_mode='sklearn', early_stop=True)
        lane0.prep_e2e()
        lane0.run_e2e(**fo)
        f_lat = [timed_step(lambda ln: ln.run_e2e(**fo), [lane0], Lane.prep_e2e) for _ in range(2)]
        f_all = [timed_step(lambda ln: ln.run_e2e(**fo), lanes, Lane.prep_e2e) for _ in range(2)]
        lane0.prep_e2e()
        fres = lane0.run_e2e(**fo)
        faithful = {"mode": "tie_mode='sklearn' (the reference's KD-tree neighbour order), early stop on: find_blobs' "
                            "default; end to end with host buffers",
                    "latency_ms_single_run": float(np.mean(f_lat)),
                    "value": B / (float(np.mean(f_all)) / 1e3), "unit": "runs/s (this rank, %d lanes)" % B,
                    "iterations": int(fres["_summary"]["n_iterations"]),
                    "n_particles_final": int(fres["_summary"]["n_particles"]),
                    "dice_2": float(fres["dice_coeff_2"])}

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        bytes_f64 = 25.0 * H * W                       # 1 B u8 in + 8 B R_b + 16 B grad out per pixel
        bytes_f32 = 13.0 * H * W                       # 1 B u8 in + 4 B R_b + 8 B grad out per pixel
        ach = bytes_f64 / (k1_ms / 1e3) / 1e9
        line = {
            "metric": WL['metric'], "value": value, "unit": "runs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(), "batch": B,
                       "tie_order": "canonical (distance, index) order: the all-GPU mode, pinned by the oracle's canonical "
                                    "variant (same order in the CPU arm); the as-is reference order (sklearn KD tree) "
                                    "is the `faithful` key",
                       "step": "one step = %d frames run concurrently (one CUDA stream + context each); the K timed "
                               "steps run back to back inside one barrier + synchronize bracket (every lane runs its "
                               "K frames one after the other); `stepwise` = the same K steps with a barrier and an L2 "
                               "flush after every step (the timing of rounds 1 / 2)" % B,
                       "padded_shape": [H, W], "n_gt": lane0.n_gt,
                       "n_particles_final": int(summ.n_particles), "n_seeds": int(summ.n_seeds),
                       "iterations": int(summ.n_iterations), "thinning_rounds": int(summ.n_thinning_rounds),
                       "l2": "inputs larger than L2: %d lanes x ~250 MB written per run >> 126 MB, so every run reads its "
                             "frame from HBM (e2e: from the host); flushed before the region and between the "
                             "`stepwise` steps (256 MiB write, untimed)" % B,
                       "parallelism": "one frame per lane, %d lanes per rank, NCCL all-gather of metric records" % B,
                       "host": {"cores": cores, "wait_policy": {"": "spin (default)", "1": "spin", "2": "yield",
                                                                 "3": "blocking"}.get(policy, policy),
                                "cores_of_this_rank": (len(os.sched_getaffinity(0))
                                                       if hasattr(os, "sched_getaffinity") else cores)}},
            "clocks": clocks,
            "stepwise": {"value": value_stepwise, "e2e": e2e_stepwise, "unit": "runs/s",
                         "ms_per_step": stepwise_ms / args.steps,
                         "what": "barrier + L2 flush after every step: the ramp (host->device copies of all lanes "
                                 "queue up before the first kernel) and the tail (the last lane finishes alone) of "
                                 "every step are inside"},
            "latency_ms_single_run": latency_ms,
            "e2e": {"value": e2e_value, "unit": "runs/s", "h2d_bytes_per_step": int(lane0.h2d * B),
                    "d2h_bytes_per_step": int(d2h * B), "latency_ms_single_run": float(np.mean(e2e_lat))},
            "gpu_launches": int(launches),
            "roofline": {
                "kernel": ("blobness_dtile_kernel" if len(WL['scales']) == 1 else "blobness_kernel (generic tile kernel, "
                           "%d scales)" % len(WL['scales'])) +
                          " (K1 f64: fused Gaussian / Hessian / eigenvalues / blobness / gradient; "
                          "the kernel the timed step launches; 25 B/px = 1 B u8 in + 8 B R_b + 16 B grad out)",
                "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": k1_dram_traffic("f64"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_f64, "ms_per_launch": k1_ms,
                "share_of_step": k1_ms / latency_ms,
                "timed": "CUDA events on the launch stream around the kernel INSIDE the timed single-lane step",
                "ms_per_launch_back_to_back": k1_f64_stream_ms,
                "frac_back_to_back": bytes_f64 / (k1_f64_stream_ms / 1e3) / 1e9 / peak,
                "measured_limiter": "FP64 pipe (exp / division / square root in FP64 per pixel), see profiles/"},
            "roofline_f32": {
                "kernel": "blobness_pair_kernel (K1 in f32 storage, the form north_star states the field tolerance "
                          "for: 13 B/px); NOT in the timed step (f32 fields flip discrete decisions of the run)",
                "bound": "hbm", "achieved": bytes_f32 / (k1_f32_stream_ms / 1e3) / 1e9, "peak": peak,
                "unit": "GB/s", "frac": bytes_f32 / (k1_f32_stream_ms / 1e3) / 1e9 / peak,
                "traffic": k1_dram_traffic("f32"),
                "algorithmic_bytes_per_launch": bytes_f32, "ms_per_launch": k1_f32_stream_ms,
                "ms_single_launch_l2_flushed": k1_f32_ms,
                "ms_per_frame_two_streams": k1_f32_two_ms,
                "frac_two_streams": bytes_f32 / (k1_f32_two_ms / 1e3) / 1e9 / peak,
                "frac_single_launch": bytes_f32 / (k1_f32_ms / 1e3) / 1e9 / peak,
                "timed": "CUDA events on the launch stream around 32 back-to-back launches over 8 distinct "
                         "frames (working set 450 MB > L2), GPU parked while the host enqueues"},
            "blobness_mpix_s": {"f32": H * W / (k1_f32_stream_ms / 1e3) / 1e6, "f64": H * W / (k1_ms / 1e3) / 1e6,
                                "f32_single_launch": H * W / (k1_f32_ms / 1e3) / 1e6,
                                "f32_ms": k1_f32_stream_ms, "f32_single_launch_ms": k1_f32_ms},
            "stage_ms_single_run": dict(zip(['dist_energy_calc', 'moving_particles', 'deleting_particles',
                                             'fix_resample', 'metrics_logging'], [float(x) for x in stage_ms])),
            "last_run": {"dice_1": float(gathered[0][0][3]), "dice_2": float(gathered[0][0][4])},
        }
        if faithful is not None:
            line["faithful"] = faithful
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(processes=1, side_rows=True)
        print(json.dumps(line))
    pool.shutdown()
    if world > 1:
        dist.destroy_process_group()
    return 0


# --------------------------------------------------------------------------------------- cfg4 / cfg5
def run_units(args):
    """BASELINE.json configs[3] / [4] through the product's sharded runner (sweep.UnitRunner): units are
    partitioned over the ranks, run on `lanes` concurrent contexts per GPU, and the per-unit records are all-gathered
    once.  One step = the whole job (host frames -> records), timed wall clock between barriers on every rank
    (device work is bracketed by the final synchronisation of each unit); value = units / max-over-ranks time."""
    import torch
    import torch.distributed as dist
    from detecting_cones_aoslo_b200 import _lib, configs, sweep
    from detecting_cones_aoslo_b200.synth import synth_aoslo, synth_config
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps({"impl": "reference", "unavailable": "the CPU arm is defined for the run workloads "
                              "(default / cfg2 / cfg3); cfg4 / cfg5 are multiples of them"}))
        return 0
    pin_rank_to_core_slice(local, world)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    n_units = args.units or WL['units']
    lanes = args.batch if args.batch > 0 else max(1, min(WL['lanes'], os.cpu_count() or 1))
    runner = sweep.UnitRunner(device=local, lanes=lanes)
    mine = list(sweep.shard_units(n_units, rank, world))
    if args.workload == "batch":
        frames_all = None
        frames = []
        for i in mine:
            im, g = synth_aoslo(SIZE, SIZE, PITCH, seed=2022 + i)
            frames.append((i, im, g, synth_config(SIZE, SIZE, PITCH, boundary=BOUNDARY, epoch_num=EPOCHS,
                                                  tie_mode=TIE_MODE, collect_stage_times=False)))
        h2d = sum(f[1].nbytes + f[2].nbytes for f in frames)

        def job():
            return runner.run_frames(frames)
    else:
        img, gt, _ = make_workload(0)
        base = synth_config(SIZE, SIZE, PITCH, boundary=BOUNDARY, tie_mode=TIE_MODE, collect_stage_times=False)
        trials = []
        for i in mine:
            c = sweep.sample_trial(base, i)
            c['region'] = dict(base['region'])             # the synthetic frame, not the shipped crop
            c['tie_mode'] = TIE_MODE
            c['collect_stage_times'] = False
            trials.append((i, c))
        h2d = img.nbytes + gt.nbytes

        def job():
            return runner.run_trials(img, gt, trials)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(1, min(args.warmup, 2))):
        job()                                               # contexts, graphs, allocator warm
    steps = max(1, min(args.steps, 3))
    barrier()
    t_timed = time.perf_counter()
    l0 = _lib.lib.aoslo_launch_count()
    rows = None
    t0 = time.perf_counter()
    for _ in range(steps):
        rows = job()
        table = sweep.gather_records(rows, n_units, device=dev)       # the one exchange: all-gather of the records
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    launches = _lib.lib.aoslo_launch_count() - l0
    barrier()
    clocks = sampler.stop(t_timed)
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    if rank == 0:
        ok = int((table[:, 1] == 0).sum())
        d2 = table[:, sweep.RECORD_FIELDS.index('dice_coeff_2')]
        value = n_units * steps / dt
        line = {"metric": WL['metric'], "value": value, "unit": "units/s", "n_gpus": world, "steps": steps,
                "warmup": max(1, min(args.warmup, 2)), "ms_per_step": dt / steps * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "%s: %d units of %s" % (args.workload, n_units, workload_name()),
                           "lanes_per_rank": lanes, "units_ok": ok, "units_failed": int(n_units - ok),
                           "early_stop": "on (reference behaviour)",
                           "l2": "inputs larger than L2 (every unit streams its own 2048^2 / 512^2 fields)",
                           "field": "computed once per rank and shared by its trials" if args.workload == "sweep"
                                    else "computed per frame"},
                "clocks": clocks,
                "e2e": {"value": value, "unit": "units/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(len(mine) * 12 * 88),
                        "note": "the job IS end to end: host frames / configs in, record table out"},
                "gpu_launches": int(launches),
                "dice_2": {"min": float(np.nanmin(d2)), "median": float(np.nanmedian(d2)), "max": float(np.nanmax(d2))}}
        print(json.dumps(line))
    runner.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0,
                    help="frames run concurrently per step (lanes per GPU); 0 = min(8, host cores)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-faithful", action="store_true", help="skip the sklearn-order / early-stop leg")
    ap.add_argument("--workload", default="default", choices=sorted(WORKLOADS),
                    help="default = the configuration BASELINE.json's metric is quoted on; cfg2 / cfg3 / batch / sweep "
                         "= its other configs")
    ap.add_argument("--units", type=int, default=0, help="batch / sweep: number of frames / trials (0 = the config's)")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.workload in ("batch", "sweep"):
        return run_units(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
