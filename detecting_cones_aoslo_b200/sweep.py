"""Sharded batch / hyper-parameter-sweep runner (replaces the reference's wandb agent loop,
main_with_delitions.py:81-176, for the hot path).

Units (frames or sweep trials) are independent -- `find_blobs` has no cross-unit dependency -- so
they are partitioned over ranks with NO data-path collective; the single exchange is one
all-gather of fixed-size per-unit records at the end (NCCL over NVLink on the GPU box, gloo in the
CPU tests).  The blobness field of an image is computed once per rank and shared by its trials
(no swept parameter changes it: sweep_config_new.yaml fixes formula / gradient / boundary).
"""
import numpy as np

RECORD_FIELDS = ('unit', 'status', 'n_iterations', 'n_particles', 'n_seeds',
                 'tp1', 'fp1', 'fn1', 'tp2', 'fp2', 'fn2',
                 'dice_coeff_1', 'dice_coeff_2', 'best_lsa_mean', 'best_blobEn_mean')
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


def sample_trial(base_config, trial_index, seed=2022):
    """Trial `i` of the sweep: i.i.d. draw with default_rng(seed + i) (wandb's `method: bayes` is
    inherently sequential and needs the wandb service)."""
    rng = np.random.default_rng(seed + int(trial_index))
    cfg = dict(base_config)
    cfg.update({k: (dict(v) if isinstance(v, dict) else v) for k, v in SWEEP_FIXED.items()})
    for key, spec in SWEEP_SPACE.items():
        if spec[0] == 'uniform':
            cfg[key] = float(rng.uniform(spec[1], spec[2]))
        elif spec[0] == 'int_uniform':
            cfg[key] = int(rng.integers(spec[1], spec[2] + 1))
        else:
            cfg[key] = spec[1][int(rng.integers(0, len(spec[1])))]
    return cfg


def shard_units(n_units, rank, world_size):
    """Contiguous block partition: unit i -> rank i * world // n (static)."""
    lo = (n_units * rank) // world_size
    hi = (n_units * (rank + 1)) // world_size
    return range(lo, hi)


def pack_record(unit, result, summary=None, last_metrics=None):
    """Fixed-size float64 record of one unit (exact for the integer fields: |v| < 2^53)."""
    summary = summary or {}
    m = last_metrics or {'1': {}, '2': {}}
    row = [unit, summary.get('status', 0), summary.get('n_iterations', -1), summary.get('n_particles', -1),
           summary.get('n_seeds', -1),
           m['1'].get('TP', -1), m['1'].get('FP', -1), m['1'].get('FN', -1),
           m['2'].get('TP', -1), m['2'].get('FP', -1), m['2'].get('FN', -1),
           result.get('dice_coeff_1', np.nan), result.get('dice_coeff_2', np.nan),
           result.get('best_lsa_mean', np.nan), result.get('best_blobEn_mean', np.nan)]
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
    """Run `run_unit(unit) -> record` for this rank's units and gather the full table.
    A failing unit must not kill the rank (the reference's drivers catch and continue,
    main_with_delitions.py:125-132): its record carries status = -1."""
    rows = []
    for u in units:
        try:
            rows.append(run_unit(u))
        except Exception:                                   # noqa: BLE001 -- mirrors the reference
            bad = np.full(RECORD_LEN, np.nan)
            bad[0], bad[1] = u, -1
            rows.append(bad)
    return gather_records(rows, n_units, device=device)


def run_sweep_on_image(img, gt, base_config, n_trials, rank=0, world_size=1, device=None):
    """cfg5: trials of the sweep space on one image, sharded over ranks (GPU)."""
    from .pipeline_with_delitions import find_blobs

    def one(i):
        cfg = sample_trial(base_config, i)
        cfg.setdefault('tie_mode', 'canonical')
        res, _ = find_blobs(cfg, np.array(gt, dtype=np.float64, copy=True), img, None, False, False)
        return pack_record(i, res, res.get('_summary'))

    return run_sharded(shard_units(n_trials, rank, world_size), one, n_trials, device=device)


def write_results_table(results, path="benchmarking_chamfer.csv"):
    """The grid driver's output (reference main_with_delitions.py:174-176): one row per `result` dict,
    sorted by `dice_coeff_1`, written as CSV.  Device tensors / private keys are dropped."""
    import pandas as pd
    rows = [{k: v for k, v in r.items() if not k.startswith('_')} for r in results]
    table = pd.DataFrame(rows).sort_values('dice_coeff_1')
    table.to_csv(path, index=False)
    return table
