"""Sharded batch / hyper-parameter-sweep runner: replaces the reference's wandb agent loop and grid driver
(main_with_delitions.py:81-176, sweep_config_new.yaml) for the hot path.

Units (frames or sweep trials) are independent -- `find_blobs` has no cross-unit dependency -- so they are
partitioned over ranks with NO data-path collective; the single exchange is one all-gather of fixed-size per-unit
records at the end (NCCL over NVLink on the GPU box, gloo in the CPU tests).

Inside a rank (`UnitRunner`):
  * the blobness field of an image is computed ONCE per (image, region, boundary, formula, gradient, scales, dtype)
    and shared by every trial that uses it (no swept parameter changes it: sweep_config_new.yaml fixes formula /
    gradient / boundary; the reference recomputes it per trial);
  * `lanes` worker threads, each with its own context (scratch) and CUDA stream, pull units from a queue; a unit is
    one `aoslo_run` call = one cached CUDA graph launch + one synchronisation.  Trials that differ only in VALUE
    parameters share the graph (the values are uploaded per call); the structural ones (neighbour-count class,
    window, receptive field) select among a few cached graphs;
  * a failing unit does not stop the rank: its record carries status -1 (the reference's drivers catch and
    continue, main_with_delitions.py:125-132, :161-168).

    python -m detecting_cones_aoslo_b200.sweep sweep --image I.tiff --gt I.csv [--space sweep.yaml] --trials N
    python -m detecting_cones_aoslo_b200.sweep grid  --dataset DIR                  (config_zoo x every image of DIR)
(under `python -m torch.distributed.run` the units are sharded over the ranks; rank 0 writes benchmarking_chamfer.csv).
"""
import itertools
import os
import queue
import threading
import time

import numpy as np

RECORD_FIELDS = ('unit', 'status', 'n_iterations', 'n_particles', 'n_seeds',
                 'tp1', 'fp1', 'fn1', 'tp2', 'fp2', 'fn2',
                 'dice_coeff_1', 'dice_coeff_2', 'best_lsa_mean', 'best_blobEn_mean', 'time_spended')
RECORD_LEN = len(RECORD_FIELDS)

# sweep_config_new.yaml parameter space (file:line of the reference in comments)
SWEEP_SPACE = {
    'lambda_dist': ('uniform', 0.05, 2.0),              # :29-32
    'mu_dist_func': ('uniform', 5.5, 6.2),              # :39-42
    'sigma_dist_func': ('uniform', 0.415, 1.66),        # :55-58
    'lambda_dist_func': ('uniform', -0.2, -0.05),       # :67-70
    'dist_n_neighbours': ('int_uniform', 1, 6),         # :71-74
    'init_blob_threshold': ('uniform', 4.95, 5.01),     # :75-78
    'threshhold_dist_del': ('uniform', 2.7, 4.5),       # :81-84
    # :85-97 lists floats (5., 5.5, ... 9); a float window raises TypeError in the reference
    # (range() step, utils.py:341), so the sampler draws from the integer members only
    'particle_call_window': ('choice', (5, 6, 7, 8, 9)),
    'fix_particles_frequency': ('int_uniform', 4, 8),   # :100-103
    'particle_call_threshold': ('uniform', 0.05, 0.15), # :104-107
}
SWEEP_FIXED = {'lambda_blob': 1.0, 'epoch_num': 30, 'metric_measure_freq': 3, 'reception_field_size': 19,
               'step_mode': 'discrete', 'blobness_formula': 'custom', 'gradient_type': 'np_grad',
               'dist_energy_type': 'pit', 'metric_algo': 'Chamfer', 'random_seed': 2022,
               'boundary_size': {'x': 15, 'y': 15}}


# ------------------------------------------------------------------------------------------ parameter space
def load_sweep_space(path, integer_windows=True):
    """Reads a wandb sweep YAML (the format of the reference's sweep_config_new.yaml) -> (space, fixed).
    `value:` entries and single-valued categoricals are fixed; `distribution: uniform | int_uniform` with min / max
    and multi-valued categoricals are sampled.  Dotted keys (`region.x_max`, `boundary_size.x`) become the nested
    dicts `find_blobs` reads.  integer_windows: drop non-integer `particle_call_window` values, which make the
    reference raise TypeError (utils.py:341) -- False keeps them (those trials then fail like the reference's)."""
    import yaml
    with open(path) as f:
        doc = yaml.safe_load(f)
    space, fixed = {}, {}

    def put_fixed(key, val):
        if '.' in key:
            outer, inner = key.split('.', 1)
            fixed.setdefault(outer, {})[inner] = val
        else:
            fixed[key] = val

    for key, spec in (doc.get('parameters') or {}).items():
        if 'value' in spec:
            put_fixed(key, spec['value'])
            continue
        dist = spec.get('distribution', 'categorical' if 'values' in spec else 'uniform')
        if 'values' in spec:
            vals = list(spec['values'])
            if key == 'particle_call_window':
                if integer_windows:
                    vals = [int(v) for v in vals if float(v) == int(v)]
                else:
                    vals = [int(v) if float(v) == int(v) else float(v) for v in vals]
            if len(vals) == 1:
                put_fixed(key, vals[0])
            else:
                space[key] = ('choice', tuple(vals))
        elif dist in ('int_uniform', 'q_uniform'):
            space[key] = ('int_uniform', int(spec['min']), int(spec['max']))
        elif dist == 'uniform':
            space[key] = ('uniform', float(spec['min']), float(spec['max']))
        else:
            raise NotImplementedError("sweep distribution %r of %r" % (dist, key))
    for flag in ('write_gif', 'verbose_func', 'save_best_result_visualisation'):
        if flag in fixed:
            fixed[flag] = bool(fixed[flag])
    return space, fixed


def sample_trial(base_config, trial_index, seed=2022, space=None, fixed=None):
    """Trial `i` of the sweep: i.i.d. draw with default_rng(seed + i), parameters in sorted key order (wandb's
    `method: bayes` is inherently sequential and needs the wandb service)."""
    space = SWEEP_SPACE if space is None else space
    fixed = SWEEP_FIXED if fixed is None else fixed
    rng = np.random.default_rng(seed + int(trial_index))
    cfg = dict(base_config)
    for k, v in fixed.items():
        cfg[k] = dict(cfg.get(k, {}), **v) if isinstance(v, dict) and isinstance(cfg.get(k), dict) else \
            (dict(v) if isinstance(v, dict) else v)
    for key, spec in space.items():
        if spec[0] == 'uniform':
            cfg[key] = float(rng.uniform(spec[1], spec[2]))
        elif spec[0] == 'int_uniform':
            cfg[key] = int(rng.integers(spec[1], spec[2] + 1))
        else:
            cfg[key] = spec[1][int(rng.integers(0, len(spec[1])))]
    return cfg


def grid_parameters(parameters):
    """reference main_with_delitions.py:17-19: the cartesian product of the list-valued entries of a config zoo."""
    keys = list(parameters)
    for values in itertools.product(*[v if isinstance(v, list) else [v] for v in parameters.values()]):
        yield dict(zip(keys, values))


def estimate_num_variant(parameters):
    """reference main_with_delitions.py:21-26."""
    n = 1
    for v in parameters.values():
        if isinstance(v, list):
            n *= len(v)
    return n


def shard_units(n_units, rank, world_size):
    """Contiguous block partition: unit i -> rank i * world // n (static)."""
    lo = (n_units * rank) // world_size
    hi = (n_units * (rank + 1)) // world_size
    return range(lo, hi)


# ------------------------------------------------------------------------------------------ records
def pack_record(unit, result, summary=None, last_metrics=None, seconds=np.nan):
    """Fixed-size float64 record of one unit (exact for the integer fields: |v| < 2^53)."""
    summary = summary or {}
    m = last_metrics or {'1': {}, '2': {}}
    row = [unit, summary.get('status', 0), summary.get('n_iterations', -1), summary.get('n_particles', -1),
           summary.get('n_seeds', -1),
           m['1'].get('TP', -1), m['1'].get('FP', -1), m['1'].get('FN', -1),
           m['2'].get('TP', -1), m['2'].get('FP', -1), m['2'].get('FN', -1),
           result.get('dice_coeff_1', np.nan), result.get('dice_coeff_2', np.nan),
           result.get('best_lsa_mean', np.nan), result.get('best_blobEn_mean', np.nan), seconds]
    return np.asarray(row, dtype=np.float64)


def failed_record(unit):
    bad = np.full(RECORD_LEN, np.nan)
    bad[0], bad[1] = unit, -1
    return bad


def _record_of(unit, result, seconds):
    last = None
    recs = result.get('_records')
    if recs:
        r = recs[-1]
        last = {'1': {'TP': r.tp[0], 'FP': r.fp[0], 'FN': r.fn[0]}, '2': {'TP': r.tp[1], 'FP': r.fp[1], 'FN': r.fn[1]}}
    return pack_record(unit, result, result.get('_summary'), last, seconds)


def _record_of_trial(unit, tr, seconds):
    """aoslo_trial_result (C ABI, aoslo_sweep) -> record row."""
    if tr.error != 0:
        return failed_record(unit)
    s = tr.summary
    if s.status != 0:                       # a device-side condition (escaped particle, NaN, capacity, ...): the
        return failed_record(unit)          # reference's find_blobs would have raised
    row = [unit, 0, s.n_iterations, s.n_particles, s.n_seeds,
           tr.last.tp[0], tr.last.fp[0], tr.last.fn[0], tr.last.tp[1], tr.last.fp[1], tr.last.fn[1],
           tr.dice_coeff_1, tr.dice_coeff_2, tr.best_lsa_mean, tr.best_blobEn_mean, seconds]
    return np.asarray(row, dtype=np.float64)


def gather_records(local_records, n_units, device=None):
    """All-gather the per-unit records of every rank -> (n_units, RECORD_LEN) table ordered by unit.
    Equal counts are required by all_gather: shards are padded with NaN rows to the largest shard."""
    import torch
    import torch.distributed as dist
    local = np.asarray(local_records, dtype=np.float64).reshape(-1, RECORD_LEN)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        table = local
    else:
        world = dist.get_world_size()
        per = max(len(shard_units(n_units, r, world)) for r in range(world))
        buf = np.full((per, RECORD_LEN), np.nan)
        buf[:local.shape[0]] = local
        t = torch.as_tensor(buf)
        if device is not None:
            t = t.to(device)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        table = np.concatenate([o.cpu().numpy() for o in out], axis=0)
        table = table[~np.isnan(table[:, 0])]
    order = np.argsort(table[:, 0], kind='stable')
    return table[order]


def run_sharded(units, run_unit, n_units, device=None):
    """Run `run_unit(unit) -> record` sequentially for this rank's units and gather the full table (host-side
    reference implementation of the sharding logic; the GPU path is UnitRunner)."""
    rows = []
    for u in units:
        try:
            rows.append(run_unit(u))
        except Exception:                                   # noqa: BLE001 -- mirrors the reference
            rows.append(failed_record(u))
    return gather_records(rows, n_units, device=device)


# ------------------------------------------------------------------------------------------ GPU runner
def _field_key(cfg):
    r, b = cfg['region'], cfg['boundary_size']
    return (r['x_min'], r['x_max'], r['y_min'], r['y_max'], int(b['x']), int(b['y']), cfg['blobness_formula'],
            cfg['gradient_type'], tuple(cfg.get('blobness_scales', (1.0,))), cfg.get('field_dtype', 'f64'))


class UnitRunner:
    """Runs units on one GPU with `lanes` concurrent contexts (one host thread + CUDA stream each)."""

    def __init__(self, device=None, lanes=4):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("detecting_cones_aoslo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self.lanes = max(1, int(lanes))
        self._engines = [dict() for _ in range(self.lanes)]      # per lane: (H, W) -> Engine
        self._streams = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)]

    def close(self):
        for d in self._engines:
            for e in d.values():
                e.close()
            d.clear()

    def _engine(self, lane, H, W, n_gt):
        from .engine import Engine
        d = self._engines[lane]
        eng = d.get((H, W))
        if eng is None or eng.gt_cap < n_gt:
            eng = Engine(H, W, gt_capacity=max(65536, n_gt), device=self.device.index)
            d[(H, W)] = eng
        return eng

    def _pool(self, jobs, work):
        """jobs -> rows; `work(lane, job)` runs on lane threads, each inside its own stream."""
        torch = self.torch
        q = queue.Queue()
        for j in jobs:
            q.put(j)
        rows, lock = [], threading.Lock()

        def loop(lane):
            torch.cuda.set_device(self.device)
            with torch.cuda.stream(self._streams[lane]):
                while True:
                    try:
                        job = q.get_nowait()
                    except queue.Empty:
                        return
                    t0 = time.perf_counter()
                    try:
                        result = work(lane, job)
                        row = _record_of(job[0], result, time.perf_counter() - t0)
                    except Exception:                       # noqa: BLE001 -- a failing unit must not stop the rank
                        row = failed_record(job[0])
                    with lock:
                        rows.append(row)

        n = min(self.lanes, max(1, len(jobs)))
        if n == 1:
            loop(0)
        else:
            ts = [threading.Thread(target=loop, args=(i,)) for i in range(n)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        rows.sort(key=lambda r: r[0])
        return rows

    def _pool_trials(self, members, Rb, grad, d_gt, H, W, chunk=4):
        """Trials on one field: the lanes pull chunks of trials and hand each chunk to ONE aoslo_sweep call (C ABI:
        the trials of a chunk run back to back on the lane's stream, best-result bookkeeping included); trials the
        C entry point does not cover (`hangarian` metric, staged mode) go through find_blobs' Python driver."""
        from . import pipeline_with_delitions as pwd
        torch = self.torch
        q = queue.Queue()
        for i in range(0, len(members), chunk):
            q.put(members[i:i + chunk])
        rows, lock = [], threading.Lock()

        def loop(lane):
            torch.cuda.set_device(self.device)
            eng = self._engine(lane, H, W, d_gt.shape[0])
            with torch.cuda.stream(self._streams[lane]):
                while True:
                    try:
                        jobs = q.get_nowait()
                    except queue.Empty:
                        return
                    t0 = time.perf_counter()
                    out, fast, structs = [], [], []
                    for unit, cfg in jobs:
                        try:
                            if cfg['metric_algo'] != 'Chamfer' or cfg.get('engine_mode', 'fused') != 'fused' or \
                                    cfg['blobness_formula'] not in ('simple_div', 'custom'):
                                res = pwd.run_on_field(eng, cfg, Rb, grad, d_gt)
                                out.append(_record_of(unit, res, time.perf_counter() - t0))
                                continue
                            cs = pwd.make_config_struct(cfg, H, W)
                            if cs.step_mode < 0:
                                raise NotImplementedError
                            fast.append(unit)
                            structs.append(cs)
                        except Exception:                   # noqa: BLE001 -- a failing unit must not stop the rank
                            out.append(failed_record(unit))
                    if structs:
                        try:
                            trs = eng.sweep(structs, Rb, grad, d_gt)
                            dt = (time.perf_counter() - t0) / len(structs)
                            out += [_record_of_trial(u, tr, dt) for u, tr in zip(fast, trs)]
                        except Exception:                   # noqa: BLE001
                            out += [failed_record(u) for u in fast]
                    with lock:
                        rows.extend(out)

        n = min(self.lanes, max(1, q.qsize()))
        if n == 1:
            loop(0)
        else:
            ts = [threading.Thread(target=loop, args=(i,)) for i in range(n)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        rows.sort(key=lambda r: r[0])
        return rows

    # ---- cfg5: many configurations on ONE image; the field is computed once per field key
    def run_trials(self, img, gt, trials):
        """trials: list of (unit_id, cur_config).  `gt` is the raw CSV array (1-based), left untouched."""
        from . import pipeline_with_delitions as pwd
        torch = self.torch
        img = np.ascontiguousarray(img)
        groups = {}
        for unit, cfg in trials:
            try:
                groups.setdefault(_field_key(cfg), []).append((unit, cfg))
            except Exception:                               # noqa: BLE001 -- malformed config: fails as a unit
                groups.setdefault(None, []).append((unit, cfg))
        rows = []
        for key, members in groups.items():
            if key is None:
                rows += [failed_record(u) for u, _ in members]
                continue
            cfg0 = members[0][1]
            gt_minus1 = np.array(gt, dtype=np.float64, copy=True) - 1.0      # utils.py:15-17 on a private copy
            H = (cfg0['region']['y_max'] - cfg0['region']['y_min']) + 2 * int(cfg0['boundary_size']['y'])
            W = (cfg0['region']['x_max'] - cfg0['region']['x_min']) + 2 * int(cfg0['boundary_size']['x'])
            master = self._engine(0, H, W, gt_minus1.shape[0])
            with torch.cuda.stream(self._streams[0]):
                d_img, d_gt = master.prepare_inputs(img, gt_minus1, cfg0['region'], cfg0['boundary_size'])
                fdt = torch.float64 if cfg0.get('field_dtype', 'f64') == 'f64' else torch.float32
                Rb, grad = master.blobness(d_img, cfg0['blobness_formula'], cfg0['gradient_type'],
                                           tuple(cfg0.get('blobness_scales', (1.0,))), fdt,
                                           out=master.field_buffers(fdt))
                master.raise_on_status()
            self._streams[0].synchronize()                   # the lanes read Rb / grad / d_gt from here on

            rows += self._pool_trials(members, Rb, grad, d_gt, H, W)
        rows.sort(key=lambda r: r[0])
        return rows

    # ---- cfg4: a batch of frames, one whole find_blobs per unit
    def run_frames(self, frames):
        """frames: list of (unit_id, img, gt, cur_config)."""
        from .pipeline_with_delitions import find_blobs

        def work(lane, job):
            unit, img, gt, cfg = job
            r, b = cfg['region'], cfg['boundary_size']
            H = (r['y_max'] - r['y_min']) + 2 * int(b['y'])
            W = (r['x_max'] - r['x_min']) + 2 * int(b['x'])
            eng = self._engine(lane, H, W, np.asarray(gt).shape[0])
            res, _ = find_blobs(cfg, np.array(gt, dtype=np.float64, copy=True), img, None, False, False, engine=eng)
            return res

        return self._pool(list(frames), work)


def run_sweep_on_image(img, gt, base_config, n_trials, rank=0, world_size=1, device=None, lanes=4, space=None,
                       fixed=None, runner=None):
    """cfg5: `n_trials` draws of the sweep space on one image, sharded over ranks, lanes inside the rank; returns the
    gathered (n_trials, RECORD_LEN) table (identical on every rank)."""
    own = runner is None
    runner = runner or UnitRunner(device=getattr(device, 'index', device), lanes=lanes)
    trials = []
    for i in shard_units(n_trials, rank, world_size):
        cfg = sample_trial(base_config, i, space=space, fixed=fixed)
        cfg.setdefault('tie_mode', 'canonical')
        cfg.setdefault('collect_stage_times', False)
        trials.append((i, cfg))
    rows = runner.run_trials(img, gt, trials)
    if own:
        runner.close()
    return gather_records(rows, n_trials, device=device)


def run_batch_of_frames(frames, rank=0, world_size=1, device=None, lanes=4, runner=None):
    """cfg4: frames = list of (img, gt, cur_config); unit i is frame i; sharded by image."""
    own = runner is None
    runner = runner or UnitRunner(device=getattr(device, 'index', device), lanes=lanes)
    mine = [(i,) + tuple(frames[i]) for i in shard_units(len(frames), rank, world_size)]
    rows = runner.run_frames(mine)
    if own:
        runner.close()
    return gather_records(rows, len(frames), device=device)


# ------------------------------------------------------------------------------------------ results table
def write_results_table(results, path="benchmarking_chamfer.csv"):
    """The grid driver's output (reference main_with_delitions.py:174-176): one row per `result` dict,
    sorted by `dice_coeff_1`, written as CSV.  Device tensors / private keys are dropped."""
    import pandas as pd
    rows = [{k: v for k, v in r.items() if not k.startswith('_')} for r in results]
    table = pd.DataFrame(rows).sort_values('dice_coeff_1')
    table.to_csv(path, index=False)
    return table


def table_to_results(table, configs):
    """Gathered record table + the per-unit configs -> `result` dicts with the reference's keys."""
    out = []
    for row in table:
        u = int(row[0])
        r = {'best_lsa_mean': row[RECORD_FIELDS.index('best_lsa_mean')],
             'best_blobEn_mean': row[RECORD_FIELDS.index('best_blobEn_mean')],
             'dice_coeff_1': row[RECORD_FIELDS.index('dice_coeff_1')],
             'dice_coeff_2': row[RECORD_FIELDS.index('dice_coeff_2')],
             'cur_config': configs[u], 'time_spended': row[RECORD_FIELDS.index('time_spended')]}
        out.append(r)
    return out


# ------------------------------------------------------------------------------------------ CLI
def _dist_setup():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    return rank, world, dev


def main(argv=None):
    import argparse
    from . import configs as named
    from .dataio import load_pair
    ap = argparse.ArgumentParser(prog="python -m detecting_cones_aoslo_b200.sweep")
    sub = ap.add_subparsers(dest='mode', required=True)
    s = sub.add_parser('sweep', help="random draws of a wandb sweep space on one image (sweep_config_new.yaml)")
    s.add_argument('--image', required=True)
    s.add_argument('--gt', required=True)
    s.add_argument('--space', default=None, help="wandb sweep YAML; default: the built-in copy of the reference's space")
    s.add_argument('--trials', type=int, default=64)
    g = sub.add_parser('grid', help="config_zoo grid x every image of a directory (main_with_delitions.py:134-176)")
    g.add_argument('--dataset', required=True)
    for p in (s, g):
        p.add_argument('--lanes', type=int, default=4)
        p.add_argument('--out', default="benchmarking_chamfer.csv")
        p.add_argument('--keep-header-row', action='store_true',
                       help="use the first CSV line as a point (the reference's pd.read_csv drops it)")
    args = ap.parse_args(argv)
    rank, world, dev = _dist_setup()
    quirk = not args.keep_header_row
    if args.mode == 'sweep':
        img, gt = load_pair(args.image, args.gt, header_row_quirk=quirk)
        space, fixed = load_sweep_space(args.space) if args.space else (None, None)
        base = named.zoo_config()
        table = run_sweep_on_image(img, gt, base, args.trials, rank, world, dev, args.lanes, space, fixed)
        cfgs = {i: sample_trial(base, i, space=space, fixed=fixed) for i in range(args.trials)}
    else:
        names = sorted(f for f in os.listdir(args.dataset) if f.lower().endswith(('.tiff', '.tif')))
        zoo = {k: [v] if not isinstance(v, (dict, list)) else v for k, v in named.zoo_config().items()}
        frames, cfgs = [], {}
        for cfg in grid_parameters(zoo):
            if cfg['blobness_formula'] == 'div_corrected' and cfg['init_blob_threshold'] > 1.:
                continue                                   # reference :143-146
            if cfg['blobness_formula'] == 'custom' and cfg['init_blob_threshold'] < 1.:
                continue
            for f in names:
                stem = os.path.splitext(f)[0]
                img, gt = load_pair(os.path.join(args.dataset, f), os.path.join(args.dataset, stem + '.csv'),
                                    header_row_quirk=quirk)
                c = dict(cfg, tie_mode=cfg.get('tie_mode', 'sklearn'))
                cfgs[len(frames)] = c
                frames.append((img, gt, c))
        table = run_batch_of_frames(frames, rank, world, dev, args.lanes)
    if rank == 0:
        out = write_results_table(table_to_results(table, cfgs), args.out)
        print("%d units, %d failed -> %s" % (len(table), int((table[:, 1] != 0).sum()), args.out))
        print(out[['dice_coeff_1', 'dice_coeff_2', 'best_lsa_mean', 'time_spended']].describe().to_string())
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
