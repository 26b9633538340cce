"""Drop-in for the hot-path helpers of the reference's utils.py.

Same names, argument meaning, return layouts and error behaviour as the reference
functions (cited per function); every computation runs in a hand-written sm_100a kernel
through libaoslo_b200 (no CPU fallback: without a CUDA device the calls raise).
Arguments and results are NumPy arrays like the reference's; each call copies its inputs
to the device and its results back (`find_blobs` keeps everything device resident instead).

Extension: functions that search neighbours take `tie_mode` ('sklearn' = the reference's
KD-tree tie order, the default; 'canonical' = (distance, index) order, the cell-list path).
"""
import threading

import numpy as np
import torch

from . import _lib
from .engine import Engine

DEFAULT_TIE_MODE = 'sklearn'
_ENGINES = {}
_ENGINES_LOCK = threading.Lock()


def get_engine(H, W, n_particles=0, n_gt=0):
    """Context cache keyed by (device, thread, padded image shape); capacity = every pixel a particle.
    A context is single-stream state, so concurrent threads get their own; an engine that is too small is
    replaced in the cache but never closed here (another caller may still hold it -- it is released when the
    last reference goes away)."""
    if not torch.cuda.is_available():
        raise RuntimeError("detecting_cones_aoslo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device()
    key = (dev, threading.get_ident(), int(H), int(W))
    need_p, need_g = max(int(H) * int(W), int(n_particles)), max(65536, int(n_gt))
    with _ENGINES_LOCK:
        eng = _ENGINES.get(key)
        if eng is None or eng.cap < need_p or eng.gt_cap < need_g:
            eng = Engine(H, W, particle_capacity=need_p, gt_capacity=need_g, device=dev)
            _ENGINES[key] = eng
    return eng


def _bbox_engine(*point_sets):
    """Engine whose frame covers the given point sets (helpers called without an image)."""
    mx = my = 1.0
    n = 0
    for p in point_sets:
        p = np.asarray(p, dtype=np.float64)
        if p.size:
            if not np.isfinite(p).all() or p.min() < 0:
                raise IndexError("particle coordinates must be finite and non-negative")
            mx, my = max(mx, float(p[:, 0].max())), max(my, float(p[:, 1].max()))
        n = max(n, p.shape[0])
    W, H = int(mx) + 2, int(my) + 2
    # round up so that repeated calls share one context
    W, H = ((W + 255) // 256) * 256, ((H + 255) // 256) * 256
    return get_engine(H, W, n_particles=n, n_gt=n)


def _tie(cur_config=None, tie_mode=None):
    if tie_mode is not None:
        return tie_mode
    if cur_config is not None:
        try:
            return cur_config['tie_mode']
        except (KeyError, TypeError):
            pass
    return DEFAULT_TIE_MODE


def _pts(eng, a):
    return eng.dev(np.asarray(a, dtype=np.float64).reshape(-1, 2), torch.float64)


def _stable(eng, flags, n):
    if flags is None:
        return torch.zeros(n, dtype=torch.uint8, device=eng.device)
    return eng.dev(np.asarray(flags).astype(np.uint8), torch.uint8)


# ------------------------------------------------------------------------------ small host helpers
def np_sigmoid(x):
    """reference utils.py:11-12."""
    return 1 / (1 + np.exp(-x))


def GT_enumerate_from_zero(gt_data):
    """reference utils.py:15-17 (mutates the caller's array)."""
    gt_data -= 1.
    return gt_data


# ------------------------------------------------------------------------------ L2 field
def calc_blobness(img, cur_config=None, scales=None, dtype=np.float64, return_grad=False):
    """NEW (the reference computes this inline, pipeline_with_delitions.py:36-81): blobness field
    R_b of a padded uint8 image.  `scales` is the multi-scale extension ([1.0] == reference)."""
    cfg = cur_config or {}
    img = np.ascontiguousarray(img)
    if img.dtype != np.uint8:
        raise TypeError("img must be uint8 (img_as_float semantics of the reference's skimage call)")
    formula = cfg.get('blobness_formula', 'custom') if hasattr(cfg, 'get') else cfg['blobness_formula']
    if formula not in ('custom', 'simple_div', 'div_corrected'):
        raise NotImplementedError
    grad_type = cfg.get('gradient_type', 'np_grad') if hasattr(cfg, 'get') else cfg['gradient_type']
    if scales is None:
        scales = cfg.get('blobness_scales', (1.0,)) if hasattr(cfg, 'get') else (1.0,)
    eng = get_engine(*img.shape)
    tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
    Rb, grad = eng.blobness(eng.upload_image(img), formula, grad_type, tuple(scales), tdt,
                            want_grad=return_grad)
    eng.raise_on_status()
    if return_grad:
        return Rb.cpu().numpy(), grad.cpu().numpy()
    return Rb.cpu().numpy()


def calc_grad_field(img, grad_type):
    """reference utils.py:152-163: (H,W,2) with [...,0] = d/drow, [...,1] = -d/dcol."""
    if grad_type not in ('np_grad', 'sobel'):
        raise AttributeError('grad_type {0} is not implemented'.format(grad_type))
    img = np.ascontiguousarray(img)
    tdt = torch.float32 if img.dtype == np.float32 else torch.float64
    eng = get_engine(*img.shape)
    return eng.grad_field(eng.to_field(img, tdt), grad_type).cpu().numpy()


# ------------------------------------------------------------------------------ L3 particle system
def initialize_high_prop_location(blobness_values, init_blob_threshold, img_shape, cur_config):
    """reference utils.py:359-375: (n,2) int64 [x, y] of inner pixels with R_b > threshold, raster order."""
    Rb = np.ascontiguousarray(blobness_values)
    eng = get_engine(*Rb.shape)
    tdt = torch.float32 if Rb.dtype == np.float32 else torch.float64
    pos = eng.seed(eng.to_field(Rb, tdt), cur_config['boundary_size']['x'], cur_config['boundary_size']['y'],
                   init_blob_threshold)
    return pos.cpu().numpy().astype(np.int64)


def calc_dist_to_neares(p_from, p_to, n_neighbours=1, return_indexes=False, tie_mode=None):
    """reference utils.py:214-222: kNN of p_to rows in p_from -> distances (n_to,k) [, indexes]."""
    p_from = np.asarray(p_from, dtype=np.float64)
    p_to = np.asarray(p_to, dtype=np.float64)
    if n_neighbours == 1 and p_from.shape[0] == p_to.shape[0]:
        assert not (p_from[:, 0] == p_to[:, 0]).all()
        assert not (p_from[:, 1] == p_to[:, 1]).all()
    eng = _bbox_engine(p_from, p_to)
    dist, idx = eng.knn(_pts(eng, p_from), _pts(eng, p_to), n_neighbours, _tie(None, tie_mode))
    eng.raise_on_status()
    if return_indexes:
        return dist.cpu().numpy(), idx.cpu().numpy().astype(np.int64)
    return dist.cpu().numpy()


def del_close_particles(particles, is_stable_particle_index, R_b, cur_config, threshold_dist=2., tie_mode=None):
    """reference utils.py:245-278 -> (new_particles (m,2) f64, to_delete {idx: bool}, new_stable list)."""
    particles = np.asarray(particles, dtype=np.float64)
    Rb = np.ascontiguousarray(R_b)
    eng = get_engine(*Rb.shape, n_particles=particles.shape[0])
    tdt = torch.float32 if Rb.dtype == np.float32 else torch.float64
    pos, st, deleted = eng.delete_close(_pts(eng, particles), _stable(eng, is_stable_particle_index,
                                                                      particles.shape[0]),
                                        eng.to_field(Rb, tdt), threshold_dist, _tie(cur_config, tie_mode))
    eng.raise_on_status()
    deleted = deleted.cpu().numpy().astype(bool)
    to_delete = {i: bool(d) for i, d in enumerate(deleted)}
    if cur_config['verbose_func']:
        print("Delete #particles=", int(deleted.sum()))
    return pos.cpu().numpy(), to_delete, [int(s) for s in st.cpu().numpy()]


def get_precomputed_dist_func(cur_config):
    """reference utils.py:95-109: rfs x rfs LUT of energy_pit_func2 (utils.py:82-87)."""
    rfs = cur_config['reception_field_size']
    eng = get_engine(256, 256)
    return eng.dist_lut(rfs, cur_config['mu_dist_func'], cur_config['sigma_dist_func'],
                        cur_config['lambda_dist_func']).cpu().numpy()


def calc_gravity_field_optim(precomputed_dist_forces_field, img_shape, particles, cur_config):
    """reference utils.py:310-333: (H,W) f64 = max(-0.48, max over particles of the stamped LUT)."""
    particles = np.asarray(particles, dtype=np.float64)
    eng = get_engine(img_shape[0], img_shape[1], n_particles=particles.shape[0])
    lut = eng.dev(np.asarray(precomputed_dist_forces_field, dtype=np.float64), torch.float64)
    field = eng.gravity_field(_pts(eng, particles), lut)
    eng.raise_on_status()
    return field.cpu().numpy()


def adding_particles(particles, is_stable_particle_index, gravity_field, img_shape, cur_config):
    """reference utils.py:336-356 -> (particles (n+a,2), stable list with a zeros appended)."""
    particles = np.asarray(particles, dtype=np.float64)
    eng = get_engine(img_shape[0], img_shape[1], n_particles=particles.shape[0])
    field = eng.dev(np.asarray(gravity_field, dtype=np.float64), torch.float64)
    pos, st = eng.add_particles(field, _pts(eng, particles), _stable(eng, is_stable_particle_index,
                                                                     particles.shape[0]),
                                cur_config['boundary_size']['x'], cur_config['boundary_size']['y'],
                                cur_config['particle_call_window'], cur_config['particle_call_threshold'])
    eng.raise_on_status()
    if cur_config['verbose_func']:
        print('#Particle Added:', pos.shape[0] - particles.shape[0])
    return pos.cpu().numpy(), [int(s) for s in st.cpu().numpy()]


def calc_distances(particles, GT_particles, n_neighbours, tie_mode=None):
    """reference utils.py:26-40: kNN(k+1) of `particles` rows in `GT_particles`, column 0 dropped;
    returns (neigh_ind (n,k), top_dist_vectors (n,k,2) = GT_particles[i] - particles[nbr])."""
    particles = np.asarray(particles, dtype=np.float64)
    GT_particles = np.asarray(GT_particles, dtype=np.float64)
    eng = _bbox_engine(particles, GT_particles)
    _, idx = eng.knn(_pts(eng, GT_particles), _pts(eng, particles), n_neighbours + 1, _tie(None, tie_mode))
    eng.raise_on_status()
    neigh_ind = idx.cpu().numpy().astype(np.int64)[:, 1:]
    return neigh_ind, GT_particles[:, None, :] - particles[neigh_ind]


def calc_dist_energy_pit(particles, n_neighbours, mu, sigma, lambda_, use_optimisation=False,
                         is_stable_particle_index=None, tie_mode=None):
    """reference utils.py:112-129 -> (force (n,2), |force| (n,))."""
    particles = np.asarray(particles, dtype=np.float64)
    eng = _bbox_engine(particles)
    stable = _stable(eng, is_stable_particle_index if use_optimisation else None, particles.shape[0])
    f = eng.dist_force(_pts(eng, particles), stable, n_neighbours, 'pit', mu, sigma, lambda_,
                       _tie(None, tie_mode)).cpu().numpy()
    eng.raise_on_status()
    return f, np.linalg.norm(f, axis=1)


def calc_dist_energy_pit_optim(particles, is_stable_particle_index, n_neighbours, mu, sigma, lambda_, tie_mode=None):
    """reference utils.py:132-134."""
    return calc_dist_energy_pit(particles, n_neighbours, mu, sigma, lambda_, use_optimisation=True,
                                is_stable_particle_index=is_stable_particle_index, tie_mode=tie_mode)


def calc_dist_energy(particles, n_neighbours, alpha, sigma, tie_mode=None):
    """reference utils.py:43-58 ('exp' energy)."""
    particles = np.asarray(particles, dtype=np.float64)
    eng = _bbox_engine(particles)
    f = eng.dist_force(_pts(eng, particles), None, n_neighbours, 'exp', alpha, sigma, 0.0,
                       _tie(None, tie_mode)).cpu().numpy()
    eng.raise_on_status()
    return f, np.linalg.norm(f, axis=1)


def move_particles(particles, is_stable_particle_index, blobness_grad, force, cur_config):
    """NEW (inline in the reference, pipeline_with_delitions.py:215-245): one move step, returns the
    moved (n,2) array (the reference mutates in place)."""
    particles = np.asarray(particles, dtype=np.float64)
    g = np.ascontiguousarray(blobness_grad)
    if cur_config['step_mode'] not in ('contin', 'discrete'):
        return particles                       # the reference silently does nothing
    eng = get_engine(g.shape[0], g.shape[1], n_particles=particles.shape[0])
    tdt = torch.float32 if g.dtype == np.float32 else torch.float64
    pos = _pts(eng, particles)
    eng.move(pos, _stable(eng, is_stable_particle_index, particles.shape[0]), eng.to_field(g, tdt),
             eng.dev(np.asarray(force, dtype=np.float64), torch.float64), cur_config['lambda_blob'],
             cur_config['lambda_dist'], cur_config['step_mode'])
    eng.raise_on_status()
    return pos.cpu().numpy()


# ------------------------------------------------------------------------------ L4 metrics
def get_particles_in_image(particles, cur_config, img_shape):
    """reference utils.py:404-418: strictly inside (b, W-b) x (b, H-b) -> (particles, indices)."""
    particles = np.asarray(particles)
    by, bx = cur_config['boundary_size']['y'], cur_config['boundary_size']['x']
    x, y = particles[:, 0], particles[:, 1]
    idx = np.nonzero((bx < x) & (x < img_shape[1] - bx) & (by < y) & (y < img_shape[0] - by))[0]
    return particles[idx], idx


def calc_acc_metrics(particles, GT_particles, gt_to_particles, particles_to_gt, cur_config, img_shape):
    """reference utils.py:378-402 from precomputed nearest distances (integer counts + 3 ratios)."""
    out = {}
    _, idx = get_particles_in_image(particles, cur_config, img_shape)
    gt_to_particles = np.asarray(gt_to_particles)
    particles_to_gt = np.asarray(particles_to_gt)
    for name, thr in (('1', 1.42), ('2', 2.83)):
        TP = np.sum([particles_to_gt <= thr])
        FN = np.sum([particles_to_gt > thr])
        FP = np.sum([gt_to_particles[idx] > thr])
        out[name] = {'TP': TP, 'FP': FP, 'FN': FN, 'dice': 2. * TP / (2. * TP + FP + FN),
                     'recall': TP / (TP + FN + 0.), 'precision': TP / (TP + FP + 0.)}
    return out


_NAN = np.float64('nan')


def _ratio(num, den):
    """num / den like NumPy scalars do it (0 / 0 -> nan instead of ZeroDivisionError); the counts are >= 0, so a
    zero denominator implies a zero numerator in all three ratios below"""
    return np.float64(num / den) if den else _NAN


def metrics_from_record(rec):
    """calc_acc_metrics dict + Chamfer summary from one device metrics record (aoslo_metrics_record).
    Plain Python arithmetic on the integer counts (IEEE double, the same values NumPy's scalars give), wrapped into
    the reference's result types at the end: ten records per run are on the latency path of find_blobs."""
    out = {}
    tp, fp, fn = rec.tp, rec.fp, rec.fn
    for t, name in enumerate(('1', '2')):
        TP, FP, FN = tp[t], fp[t], fn[t]
        out[name] = {'TP': np.int64(TP), 'FP': np.int64(FP), 'FN': np.int64(FN),
                     'dice': _ratio(2. * TP, 2. * TP + FP + FN),
                     'recall': _ratio(TP, TP + FN + 0.), 'precision': _ratio(TP, TP + FP + 0.)}
    n, g = rec.n_particles, rec.n_gt
    sum_p2g, sum_g2p = rec.sum_p2g, rec.sum_g2p
    lsa_sum = sum_p2g + sum_g2p
    lsa_mean = (sum_p2g / n if n else float('nan')) + (sum_g2p / g if g else float('nan'))
    lsa_var = rec.var_p2g + rec.var_g2p
    return out, lsa_sum, lsa_mean, lsa_var, (rec.blob_energy_sum / n if n else float('nan'))


def calc_metrics(particles, GT_particles, gt_to_particles, particles_to_gt, mode='hangarian'):
    """reference utils.py:225-242.  'Chamfer' adds the moments of the two nearest-distance arrays;
    'hangarian' solves the GT x particles assignment problem on the device (aoslo_hungarian: scipy's
    distance_matrix + linear_sum_assignment restated, same assignment incl. ties) and returns the assigned
    distances (in the order of the smaller point set) with their sum / mean / var."""
    if mode == 'Chamfer':
        return (gt_to_particles, particles_to_gt), \
            np.sum(gt_to_particles) + np.sum(particles_to_gt), \
            np.mean(gt_to_particles) + np.mean(particles_to_gt), \
            np.var(gt_to_particles) + np.var(particles_to_gt)
    if mode == 'hangarian':
        particles = np.asarray(particles, dtype=np.float64)
        GT_particles = np.asarray(GT_particles, dtype=np.float64)
        eng = _bbox_engine(particles, GT_particles)
        lsa_sum, lsa_mean, lsa_var, sel, _ = eng.hungarian(_pts(eng, GT_particles), _pts(eng, particles),
                                                           want_costs=True)
        return sel.cpu().numpy(), lsa_sum, lsa_mean, lsa_var
    raise ValueError


def calc_blob_energy(R_b, particles):
    """reference utils.py:137-144: R_b[int(y)][int(x)] per particle."""
    particles = np.asarray(particles, dtype=np.float64)
    return np.asarray(R_b)[np.trunc(particles[:, 1]).astype(np.int64), np.trunc(particles[:, 0]).astype(np.int64)]


def detection_metrics(particles, GT_particles, R_b, cur_config, is_stable_particle_index=None):
    """NEW: one fused device evaluation of calc_dist_to_neares x2 + calc_acc_metrics + calc_metrics
    ('Chamfer') + calc_blob_energy; returns (metrics dict, lsa_sum, lsa_mean, lsa_var, blob mean)."""
    particles = np.asarray(particles, dtype=np.float64)
    GT_particles = np.asarray(GT_particles, dtype=np.float64)
    Rb = np.ascontiguousarray(R_b)
    eng = get_engine(*Rb.shape, n_particles=particles.shape[0], n_gt=GT_particles.shape[0])
    tdt = torch.float32 if Rb.dtype == np.float32 else torch.float64
    rec, _, _ = eng.metrics(_pts(eng, particles), _stable(eng, is_stable_particle_index, particles.shape[0]),
                            _pts(eng, GT_particles), eng.to_field(Rb, tdt), cur_config['boundary_size']['x'],
                            cur_config['boundary_size']['y'])
    eng.raise_on_status()
    return metrics_from_record(rec)


def classify_detections(particles, GT_particles, R_b, cur_config, metric='2'):
    """NEW: the index sets the reference's visualize_all draws (utils.py:174-203) -- TP_idx / FN_idx over the
    GT points and FP_idx over the particles -- computed on the device; plotting is left to the caller."""
    metric_2_dist_map = {'1': 1.42, '2': 2.83}
    particles = np.asarray(particles, dtype=np.float64)
    GT_particles = np.asarray(GT_particles, dtype=np.float64)
    Rb = np.ascontiguousarray(R_b)
    eng = get_engine(*Rb.shape, n_particles=particles.shape[0], n_gt=GT_particles.shape[0])
    tdt = torch.float32 if Rb.dtype == np.float32 else torch.float64
    pos = _pts(eng, particles)
    bx, by = cur_config['boundary_size']['x'], cur_config['boundary_size']['y']
    rec, p2g, g2p = eng.metrics(pos, _stable(eng, None, particles.shape[0]), _pts(eng, GT_particles),
                                eng.to_field(Rb, tdt), bx, by, want_dists=True)
    gt_tp, p_fp = eng.classify(g2p, p2g, pos, bx, by, metric_2_dist_map[metric])
    eng.raise_on_status()
    gt_tp = gt_tp.cpu().numpy().astype(bool)
    return np.nonzero(gt_tp)[0], np.nonzero(~gt_tp)[0], np.nonzero(p_fp.cpu().numpy())[0]
