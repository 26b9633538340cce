"""Drop-in for the reference's pipeline_with_delitions.find_blobs (pipeline_with_delitions.py:12-403).

    find_blobs(cur_config, GT_data, img, img_colorful, VERBOSE, USE_WANDB, experiment_idx=0)
        -> (result, time_spended)

Same signature, config keys, result / time_spended keys, in-place GT mutation and exceptions
as the reference.  Only the GT "-1" shift (a side effect on the caller's array) happens on the host;
np.round, crop and mirror padding, the blobness field, gradient, seeding, thinning, gravity-field
resampling, kNN pit force, move, deletion and metrics run in hand-written sm_100a kernels, device
resident, through libaoslo_b200's C ABI.  There is no CPU fallback.

Extension keys (optional, absent == reference behaviour):
    tie_mode          'sklearn' (default; neighbour ties ordered like the reference's KD tree)
                      | 'canonical' ((distance, index) order: the all-GPU cell-list path)
    field_dtype       'f64' (default) | 'f32'
    blobness_scales   [1.0] (default) -- multi-scale extension
    early_stop        True (default; reference :381-382) | False (run all epoch_num iterations)
    engine_mode       'fused' (default: one aoslo_run call) | 'staged' (stage calls from Python)
    collect_stage_times  True (default: the five `time_spended` buckets are measured) | False (zeros; saves
                      the five one-thread time-stamp kernels per iteration of the canonical run graph)
VERBOSE (matplotlib figures), write_gif and save_best_result_visualisation are visualisation
side effects outside the accelerated path and are ignored.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from .image_transformation import crop_image, mirror_padding_image
from .utils import GT_enumerate_from_zero, get_engine, metrics_from_record

TIME_KEYS = ('dist_energy_calc', 'moving_particles', 'deleting_particles', 'fix_resample', 'metrics_logging')


def _get(cfg, key, default):
    try:
        return cfg[key]
    except (KeyError, TypeError):
        return default


def make_config_struct(cur_config, H, W, metric_every_iteration=False):
    """Validate the reference's config dict once and pack it into the POD `aoslo_config`."""
    if cur_config['dist_energy_type'] not in _lib.ENERGY:
        raise AttributeError('dist_energy_type={} is not implemented!'.format(cur_config['dist_energy_type']))
    step_mode = cur_config['step_mode']
    window = cur_config['particle_call_window']
    if not isinstance(window, (int, np.integer)) or isinstance(window, bool):
        raise TypeError("'%s' object cannot be interpreted as an integer" % type(window).__name__)
    c = _lib.Config()
    c.H, c.W = int(H), int(W)
    c.bx, c.by = int(cur_config['boundary_size']['x']), int(cur_config['boundary_size']['y'])
    c.lambda_dist, c.lambda_blob = float(cur_config['lambda_dist']), float(cur_config['lambda_blob'])
    c.dist_alpha, c.dist_sigma = float(cur_config['dist_alpha']), float(cur_config['dist_sigma'])
    c.dist_n_neighbours = int(cur_config['dist_n_neighbours'])
    c.epoch_num = int(cur_config['epoch_num'])
    c.metric_measure_freq = 1 if metric_every_iteration else int(cur_config['metric_measure_freq'])
    c.step_mode = _lib.STEP.get(step_mode, -1)
    c.dist_energy_type = _lib.ENERGY[cur_config['dist_energy_type']]
    c.mu_dist_func = float(cur_config['mu_dist_func'])
    c.sigma_dist_func = float(cur_config['sigma_dist_func'])
    c.lambda_dist_func = float(cur_config['lambda_dist_func'])
    c.threshhold_dist_del = float(cur_config['threshhold_dist_del'])
    c.particle_call_window = int(window)
    c.particle_call_threshold = float(cur_config['particle_call_threshold'])
    c.init_blob_threshold = float(cur_config['init_blob_threshold'])
    c.reception_field_size = int(cur_config['reception_field_size'])
    c.fix_particles_frequency = int(cur_config['fix_particles_frequency'])
    c.tie_mode = _lib.TIE[_get(cur_config, 'tie_mode', 'sklearn')]
    c.early_stop = 1 if _get(cur_config, 'early_stop', True) else 0
    c.thinning_threshold = 3.0
    return c


class Trace:
    """Optional recorder of per-stage results (same event kinds as the oracle's trace)."""

    def __init__(self):
        self.events = []

    def add(self, kind, **payload):
        self.events.append((kind, payload))

    def of(self, kind):
        return [p for k, p in self.events if k == kind]


def _np(t):
    return t.detach().cpu().numpy()


def prepare_inputs(cur_config, GT_data, img, img_colorful=None):
    """reference pipeline_with_delitions.py:26-33 (host): GT -1 in place, round, crop, mirror pad."""
    GT_data = GT_enumerate_from_zero(GT_data)
    GT_data = np.round(GT_data)
    img, img_colorful, GT_data = crop_image(img, img_colorful, GT_data, cur_config['region'])
    img, GT_data = mirror_padding_image(img, cur_config['boundary_size'], GT_data)
    return np.ascontiguousarray(img), np.ascontiguousarray(GT_data, dtype=np.float64)


def _track(result, pm, lsa_mean, lsa_var, blob_mean, n_gt, n_particles, use_wandb, experiment_idx):
    """best-result bookkeeping of the reference (:297-317 wandb branch, :335-359 otherwise)."""
    for t, key in (('1', 'dice_coeff_1'), ('2', 'dice_coeff_2')):
        if result[key] < pm[t]['dice'] * 0.995:
            result[key] = pm[t]['dice']
            if use_wandb:
                result['lsa_mean'] = lsa_mean
                result['lsa_var'] = lsa_var
                result['#GT/#particles'] = n_gt / n_particles
                result['experiment_idx'] = experiment_idx
    if not use_wandb:
        if result['best_lsa_mean'] > lsa_mean:
            result['best_lsa_mean'] = lsa_mean
        if result['best_blobEn_mean'] < blob_mean:
            result['best_blobEn_mean'] = blob_mean


def find_blobs(cur_config, GT_data, img, img_colorful, VERBOSE, USE_WANDB, experiment_idx=0, trace=None,
               log_fn=None, engine=None):
    """`trace`, `log_fn`, `engine` are extensions: a stage recorder, a sink for the per-iteration log
    records the reference sends to wandb, and a caller-owned Engine (one per concurrent stream)."""
    result = {
        'best_lsa_mean': +np.inf,
        'best_blobEn_mean': -np.inf,
        'dice_coeff_1': -np.inf,
        'dice_coeff_2': -np.inf,
        'cur_config': cur_config,
        'time_spended': 0,
    }
    t0 = time.time()
    np.random.seed(cur_config['random_seed'])          # reference :24 (no RNG is consumed afterwards)

    region, boundary = cur_config['region'], cur_config['boundary_size']
    if not isinstance(img, torch.Tensor):                  # a device uint8 tensor (dataio.PinnedBatch) is taken as is
        img = np.asarray(img)
    assert (0 <= region['x_min']) and (0 <= region['y_min'])                            # crop_image asserts
    assert (img.shape[1] > region['x_max']) and (img.shape[0] > region['y_max'])
    H = (region['y_max'] - region['y_min']) + 2 * int(boundary['y'])
    W = (region['x_max'] - region['x_min']) + 2 * int(boundary['x'])
    formula = cur_config['blobness_formula']
    if formula not in ('simple_div', 'custom', 'div_corrected'):
        raise NotImplementedError
    field_dtype = torch.float64 if _get(cur_config, 'field_dtype', 'f64') == 'f64' else torch.float32
    scales = tuple(_get(cur_config, 'blobness_scales', (1.0,)))
    mode = _get(cur_config, 'engine_mode', 'fused')
    if trace is not None or cur_config['metric_algo'] == 'hangarian':
        mode = 'staged'          # the assignment solve is a stage of its own (aoslo_hungarian)

    eng = engine if engine is not None else get_engine(H, W, n_gt=GT_data.shape[0])
    if eng.H != H or eng.W != W or eng.gt_cap < GT_data.shape[0]:
        raise ValueError('engine was created for a different padded shape / GT capacity')
    eng.clear_status()                                  # stream ordered; the word is read once, after the run
    # reference :26-33: GT -= 1 on the CALLER's array (host side effect), then round / crop / mirror
    # padding -- on the device (aoslo_prepare_inputs).
    # fused canonical run: the cropped GT count is only needed on the device, so nothing synchronises between the
    # upload and the end of the run; the in-place shift of the caller's array (one pass over it on the host) is
    # performed by the library while the device runs (aoslo_stage_gt), the device copy is shifted by the kernel
    lazy_gt = (mode == 'fused' and _get(cur_config, 'tie_mode', 'sklearn') == 'canonical'
               and cur_config['metric_algo'] == 'Chamfer' and GT_data.shape[0] > 0)
    defer_shift = (lazy_gt and isinstance(GT_data, np.ndarray) and GT_data.dtype == np.float64
                   and GT_data.flags.c_contiguous and GT_data.flags.writeable and GT_data.ndim == 2
                   and GT_data.shape[1] == 2)
    if not defer_shift:
        GT_data = GT_enumerate_from_zero(GT_data)
    try:
        d_img, d_gt = eng.prepare_inputs(img, GT_data, region, boundary, sync=not lazy_gt, gt_is_raw=defer_shift)
        if d_gt.shape[0] < 1:
            raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by NearestNeighbors.")
        Rb, grad = eng.blobness(d_img, formula, cur_config['gradient_type'], scales, field_dtype,
                                out=eng.field_buffers(field_dtype))
        if mode != 'fused':
            eng.raise_on_status()                       # assert (eigs[0] > eigs[1]).all()  (:43)
        # fused: the assert flag stays in the status word and is raised from the run's summary (one sync per call)

        if mode == 'fused':
            time_spended = _run_fused(eng, cur_config, Rb, grad, d_gt, result, USE_WANDB, experiment_idx, log_fn,
                                      n_gt_on_device=lazy_gt)
        else:
            time_spended = _run_staged(eng, cur_config, Rb, grad, d_gt, result, USE_WANDB, experiment_idx, log_fn,
                                       trace)
    finally:
        if defer_shift:
            eng.flush_host_work()                       # error exits: the caller's array is shifted like the reference's
    result['time_spended'] = time.time() - t0
    return result, time_spended


def run_on_field(eng, cur_config, Rb, grad, d_gt, USE_WANDB=False, experiment_idx=0, log_fn=None):
    """The part of find_blobs after the field computation (reference :95-403) on an existing blobness field:
    what a sweep calls per trial, the field being shared by every trial of an image (sweep.UnitRunner).
    `Rb` / `grad` / `d_gt` are device tensors (Engine.blobness / Engine.prepare_inputs, possibly of another
    Engine of the same shape).  Returns the `result` dict of find_blobs."""
    result = {
        'best_lsa_mean': +np.inf,
        'best_blobEn_mean': -np.inf,
        'dice_coeff_1': -np.inf,
        'dice_coeff_2': -np.inf,
        'cur_config': cur_config,
        'time_spended': 0,
    }
    t0 = time.time()
    if cur_config['blobness_formula'] not in ('simple_div', 'custom', 'div_corrected'):
        raise NotImplementedError
    eng.clear_status()
    if cur_config['metric_algo'] == 'hangarian' or _get(cur_config, 'engine_mode', 'fused') == 'staged':
        _run_staged(eng, cur_config, Rb, grad, d_gt, result, USE_WANDB, experiment_idx, log_fn, None)
    else:
        _run_fused(eng, cur_config, Rb, grad, d_gt, result, USE_WANDB, experiment_idx, log_fn)
    result['time_spended'] = time.time() - t0
    return result


def _log_record(log_fn, it, pm, lsa_sum, lsa_mean, lsa_var, blob_sum, blob_mean, n_gt, n, n_stable):
    """the reference's wandb.log payload (:270-295), handed to a caller-supplied sink."""
    if log_fn is None:
        return
    logs = {'step': it, 'blob_energy_sum': blob_sum, 'blob_energy_mean': blob_mean,
            'GT_p_amount/particles_amount': n_gt / n}
    for t, m in pm.items():
        for k, v in m.items():
            logs[t + '_' + k] = v
    logs.update({'linear_sum_assignment_sum': lsa_sum, 'linear_sum_assignment_mean': lsa_mean,
                 'linear_sum_assignment_var': lsa_var, 'amount_particles': n, 'amount_stable_particles': n_stable,
                 'stable_particles_persentage': n_stable / n})
    log_fn(logs)


def _run_fused(eng, cur_config, Rb, grad, d_gt, result, use_wandb, experiment_idx, log_fn, n_gt_on_device=False):
    cfg = make_config_struct(cur_config, eng.H, eng.W, metric_every_iteration=use_wandb)
    if cfg.step_mode < 0:
        raise NotImplementedError("unknown step_mode is a silent no-op in the reference; use engine_mode='staged'")
    if cur_config['metric_algo'] != 'Chamfer':
        raise ValueError                                   # 'hangarian' runs in the staged driver (find_blobs)
    summ, recs, pos, st, ms = eng.run(cfg, Rb, grad, d_gt, collect_times=bool(_get(cur_config, 'collect_stage_times', True)),
                                      n_gt_on_device=n_gt_on_device)
    if n_gt_on_device and recs and recs[0].n_gt < 1:      # every GT point fell outside the region (sklearn raises)
        raise ValueError("Found array with 0 sample(s) while a minimum of 1 is required by NearestNeighbors.")
    if summ.status:
        _lib.raise_device_status(summ.status)
    for rec in recs:
        pm, lsa_sum, lsa_mean, lsa_var, blob_mean = metrics_from_record(rec)
        _track(result, pm, lsa_mean, lsa_var, blob_mean, rec.n_gt, rec.n_particles, use_wandb, experiment_idx)
        if use_wandb:
            _log_record(log_fn, rec.iteration, pm, lsa_sum, lsa_mean, lsa_var, rec.blob_energy_sum, blob_mean,
                        rec.n_gt, rec.n_particles, rec.n_stable)
    result['_final_particles'] = pos
    result['_final_stable'] = st
    result['_records'] = recs
    result['_summary'] = {k: getattr(summ, k) for k, _ in summ._fields_}
    return {k: ms[i] / 1e3 for i, k in enumerate(TIME_KEYS)}


def _run_staged(eng, cur_config, Rb, grad, d_gt, result, use_wandb, experiment_idx, log_fn, trace):
    """The same run driven stage by stage from Python (one C-ABI call per reference function);
    used for tracing / parity tests.  Mirrors the control flow of the reference line by line."""
    tie = _get(cur_config, 'tie_mode', 'sklearn')
    bx, by = cur_config['boundary_size']['x'], cur_config['boundary_size']['y']
    early_stop = _get(cur_config, 'early_stop', True)
    ts = {k: 0. for k in TIME_KEYS}

    lut = eng.dist_lut(cur_config['reception_field_size'], cur_config['mu_dist_func'],
                       cur_config['sigma_dist_func'], cur_config['lambda_dist_func'])
    if trace is not None:
        trace.add('lut', lut=_np(lut))
        trace.add('grad', R_b=_np(Rb), grad=_np(grad))
    pos = eng.seed(Rb, bx, by, cur_config['init_blob_threshold'])
    if trace is not None:
        trace.add('seeds', particles=_np(pos).astype(np.int64))
    stable = torch.zeros(pos.shape[0], dtype=torch.uint8, device=eng.device)

    def delete(pos, stable, thr):
        new_pos, new_st, deleted = eng.delete_close(pos, stable, Rb, thr, tie)
        if trace is not None:
            trace.add('delete', particles_in=_np(pos), stable_in=_np(stable).astype(np.int8), threshold=float(thr),
                      particles_out=_np(new_pos), deleted=_np(deleted).astype(bool),
                      stable_out=_np(new_st).astype(np.int8))
        return new_pos, new_st, int(pos.shape[0] - new_pos.shape[0])

    while True:
        pos, stable, ndel = delete(pos, stable, 3)
        if ndel == 0:
            break
    stable = torch.ones_like(stable)

    def resample(pos, stable):
        field = eng.gravity_field(pos, lut)
        if trace is not None:
            trace.add('gravity', particles=_np(pos), field=_np(field))
        pos, stable = eng.add_particles(field, pos, stable, bx, by, cur_config['particle_call_window'],
                                        cur_config['particle_call_threshold'])
        if trace is not None:
            trace.add('adding', particles_out=_np(pos), stable_out=_np(stable).astype(np.int8))
        return pos, stable

    pos, stable = resample(pos, stable)

    def sync_time(t1, key):
        torch.cuda.synchronize(eng.device)
        ts[key] += time.time() - t1
        return time.time()

    for iteration in range(cur_config['epoch_num']):
        t1 = time.time()
        k = cur_config['dist_n_neighbours']
        if cur_config['dist_energy_type'] == 'exp':
            force = eng.dist_force(pos, None, k, 'exp', cur_config['dist_alpha'], cur_config['dist_sigma'], 0.0, tie)
            if trace is not None:
                trace.add('force_exp', particles=_np(pos), force=_np(force))
        elif cur_config['dist_energy_type'] == 'pit':
            force = eng.dist_force(pos, stable, k, 'pit', cur_config['mu_dist_func'],
                                   cur_config['sigma_dist_func'], cur_config['lambda_dist_func'], tie)
            if trace is not None:
                trace.add('force', particles=_np(pos), stable=_np(stable).astype(np.int8), force=_np(force))
        else:
            raise AttributeError('dist_energy_type={} is not implemented!'.format(cur_config['dist_energy_type']))
        t1 = sync_time(t1, 'dist_energy_calc')
        if cur_config['step_mode'] in ('contin', 'discrete'):
            pos = pos.clone()
            eng.move(pos, stable, grad, force, cur_config['lambda_blob'], cur_config['lambda_dist'],
                     cur_config['step_mode'])
        t1 = sync_time(t1, 'moving_particles')
        pos, stable, _ = delete(pos, stable, cur_config['threshhold_dist_del'])
        t1 = sync_time(t1, 'deleting_particles')
        if iteration % cur_config['fix_particles_frequency'] == 3:
            stable = torch.ones_like(stable)
            pos, stable = resample(pos, stable)
        t1 = sync_time(t1, 'fix_resample')
        if use_wandb or iteration % cur_config['metric_measure_freq'] == 0:
            if cur_config['metric_algo'] not in ('Chamfer', 'hangarian'):
                raise ValueError
            rec, p2g, g2p = eng.metrics(pos, stable, d_gt, Rb, bx, by, want_dists=trace is not None)
            pm, lsa_sum, lsa_mean, lsa_var, blob_mean = metrics_from_record(rec)
            if cur_config['metric_algo'] == 'hangarian':       # calc_metrics(mode='hangarian'), utils.py:228-234
                lsa_sum, lsa_mean, lsa_var, _, _ = eng.hungarian(d_gt, pos)
                eng.raise_on_status()
            _track(result, pm, lsa_mean, lsa_var, blob_mean, rec.n_gt, rec.n_particles, use_wandb, experiment_idx)
            if use_wandb:
                _log_record(log_fn, iteration, pm, lsa_sum, lsa_mean, lsa_var, rec.blob_energy_sum, blob_mean,
                            rec.n_gt, rec.n_particles, rec.n_stable)
            if trace is not None:
                pn = _np(pos)
                rb = _np(Rb)
                trace.add('blob_energy', energy=rb[np.trunc(pn[:, 1]).astype(np.int64),
                                                   np.trunc(pn[:, 0]).astype(np.int64)].astype(np.float64))
                trace.add('acc', particles=pn, gt=_np(d_gt), gt_to_particles=_np(p2g)[:, None],
                          particles_to_gt=_np(g2p)[:, None], metrics=pm)
                trace.add('chamfer', lsa_sum=float(lsa_sum), lsa_mean=float(lsa_mean), lsa_var=float(lsa_var))
        eng.raise_on_status()
        n_stable = int(stable.sum().item())
        if early_stop and n_stable / pos.shape[0] > 0.997:
            break
        sync_time(t1, 'metrics_logging')
    result['_final_particles'] = pos
    result['_final_stable'] = stable
    return ts
