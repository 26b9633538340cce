"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the detection hot path.

A NumPy / SciPy / scikit-learn restatement of the reference's `find_blobs`
path (EgorovEV/Detecting_cones_AOSLO).  Every function cites the reference
file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import this module, and only as the
checker / the timed CPU arm -- the product package never imports it.

Pinning: the reference ships no tests or golden vectors.  This oracle is pinned
against outputs of the reference ITSELF, executed in the build container through
oracle/reference_harness.py (unmodified reference modules + 3 stubs) and
committed under tests/golden/ by oracle/gen_golden.py; tests/test_oracle_golden.py
replays them.  The scikit-image boundary (Hessian) is restated from its
published algorithm (oracle/skimage_shim.py) -- that boundary is "parity
unpinned" by the reference.

Where the reference loops in Python over particles the oracle uses whole-array
NumPy expressions ONLY when the arithmetic (operation order per element) is
unchanged, so results are bit-identical; the sequential greedy deletion stays a
loop because its result depends on row order.

kNN tie order: `tie_mode='sklearn'` calls the very same sklearn
NearestNeighbors(kd_tree, leaf_size=1) the reference calls (utils.py:29, :218);
`tie_mode='canonical'` orders neighbours by (squared distance, index), which is
what the GPU cell-list kNN produces.
"""
import time

import numpy as np
from scipy import ndimage
from scipy.optimize import linear_sum_assignment
from scipy.spatial import cKDTree, distance_matrix
from sklearn.neighbors import NearestNeighbors

from oracle import skimage_shim


# --------------------------------------------------------------------------- L1 pre-processing
def GT_enumerate_from_zero(gt_data):
    """utils.py:15-17 (mutates the caller's array)."""
    gt_data -= 1.
    return gt_data


def crop_image(img, img_colorful, gt_particles, region):
    """image_transformation.py:6-21."""
    assert (0 <= region['x_min']) and (0 <= region['y_min'])
    assert (img.shape[1] > region['x_max']) and (img.shape[0] > region['y_max'])
    img_crop = img[region['y_min']:region['y_max'], region['x_min']:region['x_max']]
    col_crop = None
    if img_colorful is not None:
        col_crop = img_colorful[region['y_min']:region['y_max'], region['x_min']:region['x_max'], :]
    keep_y = np.logical_and(gt_particles[:, 1] > region['y_min'], gt_particles[:, 1] < region['y_max'])
    keep_x = np.logical_and(gt_particles[:, 0] > region['x_min'], gt_particles[:, 0] < region['x_max'])
    crop_particles = gt_particles[np.logical_and(keep_y, keep_x)]
    crop_particles[:, 0] -= region['x_min']
    crop_particles[:, 1] -= region['y_min']
    return img_crop, col_crop, crop_particles


def mirror_padding_image(img, boundary_size, GT_data):
    """image_transformation.py:23-35 (== np.pad(img, b, 'symmetric'); GT shifted in place)."""
    by, bx = boundary_size['y'], boundary_size['x']
    padded = np.pad(img, ((by, by), (bx, bx)), mode='symmetric')
    GT_data[:, 1] += by
    GT_data[:, 0] += bx
    return padded, GT_data


# --------------------------------------------------------------------------- L2 field
def np_sigmoid(x):
    """utils.py:11-12."""
    return 1 / (1 + np.exp(-x))


def hessian_eigs(img, sigma=1.0):
    """pipeline_with_delitions.py:36-38 through the restated skimage functions.
    For sigma != 1 (multi-scale extension, not in the reference) the Hessian is
    multiplied by sigma^2 (scale normalisation); sigma == 1 is exactly the reference."""
    H = skimage_shim.hessian_matrix(img, sigma=sigma, mode='constant', cval=0, order='rc')
    if sigma != 1:
        s2 = float(sigma) * float(sigma)
        H = [h * s2 for h in H]
    return skimage_shim.hessian_matrix_eigvals(H)


def blobness_from_eigs(eigs, formula):
    """pipeline_with_delitions.py:60-81.  Only 'custom' is exercised by the
    reference's drivers; 'simple_div' is restated, 'div_corrected' raises the
    IndexError the reference raises on numpy >= 1.23."""
    if formula == 'custom':
        return eigs[1, :, :] - eigs[0, :, :] + (np_sigmoid(-eigs[0, :, :]) * 10.)
    if formula == 'simple_div':
        R_b = np.divide(np.abs(eigs[0, :, :]), np.abs(eigs[1, :, :]))
        threshold = 10.
        R_b[R_b > threshold] = threshold
        return R_b / threshold
    if formula == 'div_corrected':
        raise IndexError('div_corrected: list-of-mask indexing is broken in the reference '
                         '(pipeline_with_delitions.py:77)')
    raise NotImplementedError


def calc_blobness(img, formula='custom', scales=(1.0,)):
    """Blobness field R_b of a padded uint8 image: per-scale eigenvalues ->
    formula -> max over scales (scales=(1.0,) == the reference, :36-68)."""
    R_b = None
    for s in scales:
        eigs = hessian_eigs(img, s)
        assert (eigs[0, :, :] > eigs[1, :, :]).all()      # pipeline_with_delitions.py:43
        r = blobness_from_eigs(eigs, formula)
        R_b = r if R_b is None else np.maximum(R_b, r)
    return R_b


def calc_grad_field(img, grad_type):
    """utils.py:152-163: [...,0] = d/drow, [...,1] = -d/dcol."""
    if grad_type == 'np_grad':
        g_on_x_axis, g_on_y_axis = np.gradient(img)
    elif grad_type == 'sobel':
        g_on_x_axis = ndimage.sobel(img, axis=0, mode='constant')
        g_on_y_axis = ndimage.sobel(img, axis=1, mode='constant')
    else:
        raise AttributeError('grad_type {0} is not implemented'.format(grad_type))
    g_on_y_axis = -g_on_y_axis
    return np.stack((g_on_x_axis, g_on_y_axis), axis=2)


# --------------------------------------------------------------------------- kNN
def knn(p_from, p_to, k, tie_mode='sklearn'):
    """k nearest points of `p_from` for every row of `p_to` -> (dist, idx), both (n_to, k).
    'sklearn' == utils.py:218-219 / :29-31 verbatim."""
    p_from = np.ascontiguousarray(p_from, dtype=np.float64)
    p_to = np.ascontiguousarray(p_to, dtype=np.float64)
    if tie_mode == 'sklearn':
        x_nn = NearestNeighbors(n_neighbors=k, leaf_size=1, algorithm='kd_tree', metric='l2').fit(p_from)
        return x_nn.kneighbors(p_to)
    if tie_mode != 'canonical':
        raise ValueError(tie_mode)
    n = p_from.shape[0]
    if k > n:
        raise ValueError('Expected n_neighbors <= n_samples_fit')
    tree = cKDTree(p_from)
    kk = min(n, k + 12)
    while True:
        _, idx = tree.query(p_to, k=kk)
        idx = idx.reshape(p_to.shape[0], kk)
        diff = p_to[:, None, :] - p_from[idx]
        d2 = diff[:, :, 0] * diff[:, :, 0] + diff[:, :, 1] * diff[:, :, 1]
        order = np.lexsort((idx, d2), axis=1)
        d2s = np.take_along_axis(d2, order, axis=1)
        idxs = np.take_along_axis(idx, order, axis=1)
        # the k-th neighbour is only determined if the candidate list reaches strictly
        # beyond its distance (all ties at the boundary are inside the list)
        if kk == n or (d2s[:, kk - 1] > d2s[:, k - 1]).all():
            return np.sqrt(d2s[:, :k]), idxs[:, :k]
        kk = min(n, kk * 2)


def calc_dist_to_neares(p_from, p_to, n_neighbours=1, return_indexes=False, tie_mode='sklearn'):
    """utils.py:214-222."""
    if n_neighbours == 1 and p_from.shape[0] == p_to.shape[0]:
        assert not (p_from[:, 0] == p_to[:, 0]).all()
        assert not (p_from[:, 1] == p_to[:, 1]).all()
    min_dists, indexes = knn(p_from, p_to, n_neighbours, tie_mode)
    if return_indexes:
        return min_dists, indexes
    return min_dists


# --------------------------------------------------------------------------- L3 particle system
def _trunc_index(a):
    """int(particle[..]) of the reference: truncation toward zero."""
    return np.trunc(a).astype(np.int64)


def initialize_high_prop_location(blobness_values, init_blob_threshold, img_shape, cur_config):
    """utils.py:359-375: raster-order (x, y) of inner pixels with R_b > threshold."""
    by, bx = cur_config['boundary_size']['y'], cur_config['boundary_size']['x']
    inner = blobness_values[by:img_shape[0] - by, bx:img_shape[1] - bx]
    all_y, all_x = np.where(inner > init_blob_threshold)
    return np.vstack((all_x + bx, all_y + by)).T


def del_close_particles(particles, is_stable_particle_index, R_b, cur_config, threshold_dist=2.,
                        tie_mode='sklearn'):
    """utils.py:245-278: sequential greedy deletion in kNN-row order, then an
    order-preserving compaction.  Returns (new_particles f64, deleted bool[N], new_stable list)."""
    n = particles.shape[0]
    min_dists, indexes = calc_dist_to_neares(particles, particles, n_neighbours=2, return_indexes=True,
                                             tie_mode=tie_mode)
    stable = np.asarray(is_stable_particle_index, dtype=bool)
    deleted = np.zeros(n, dtype=bool)
    pf = np.asarray(particles, dtype=np.float64)
    iy = _trunc_index(pf[:, 1])
    ix = _trunc_index(pf[:, 0])
    rows = np.nonzero((min_dists[:, 1] < threshold_dist) & ~stable[indexes[:, 0]])[0]
    for r in rows:
        f = indexes[r, 0]
        t = indexes[r, 1]
        if not deleted[f] and not deleted[t]:
            energy1 = R_b[iy[f]][ix[f]]
            energy2 = R_b[iy[t]][ix[t]]
            deleted[f if energy1 < energy2 else t] = True
    keep = ~deleted
    new_particles = np.zeros((int(keep.sum()), 2))
    new_particles[:] = pf[keep]
    new_stable = [int(s) for s in np.asarray(is_stable_particle_index)[keep]]
    return new_particles, deleted, new_stable


def energy_pit_func2_closed(dist, mu, sigma, lambda_):
    """utils.py:82-87 with scipy.stats.expon(lambda_).pdf and scipy.stats.norm(mu, sigma).pdf
    written out in scipy's own operation order (expon: loc=lambda_, scale=1;
    norm pdf = exp(-z**2/2)/sqrt(2*pi)/sigma).  Array-valued."""
    dist = np.asarray(dist, dtype=np.float64)
    z = dist - lambda_
    with np.errstate(over='ignore', invalid='ignore'):
        e = np.where(z >= 0, np.exp(-z), 0.0)
        y = (dist - mu) / sigma
        nrm = np.exp(-y ** 2 / 2.0) / np.sqrt(2 * np.pi) / sigma
        out = np.where(dist <= mu, e - nrm + 0.48, e + nrm - 0.48)
    # NaN distance: both comparisons false in the reference -> second branch with NaNs
    return out


def get_precomputed_dist_func(cur_config):
    """utils.py:95-109: LUT[y][x] = f(||(x-h, y-h)||), h = rfs // 2."""
    rfs = cur_config['reception_field_size']
    d = np.arange(rfs, dtype=np.float64) - (rfs // 2)
    dist = np.sqrt((d[None, :] ** 2.) + (d[:, None] ** 2.))
    return energy_pit_func2_closed(dist, cur_config['mu_dist_func'], cur_config['sigma_dist_func'],
                                   cur_config['lambda_dist_func'])


def calc_gravity_field_optim(precomputed_dist_forces_field, img_shape, particles, cur_config):
    """utils.py:310-333: running max of the LUT stamped at every particle."""
    gravity_field = np.zeros(img_shape) - 0.48
    rfs = cur_config['reception_field_size']
    h = rfs // 2
    for particle in particles:
        cx, cy = int(particle[0]), int(particle[1])
        low_x, low_y = max(0, cx - h), max(0, cy - h)
        high_x, high_y = min(img_shape[1], cx + h + 1), min(img_shape[0], cy + h + 1)
        lkx, lky = -1 * min(0, cx - h), -1 * min(0, cy - h)
        hkx = rfs - max(0, (cx + h + 1 - img_shape[1]))
        hky = rfs - max(0, (cy + h + 1 - img_shape[0]))
        gravity_field[low_y:high_y, low_x:high_x] = np.maximum(
            precomputed_dist_forces_field[lky:hky, lkx:hkx], gravity_field[low_y:high_y, low_x:high_x])
    return gravity_field


def adding_particles(particles, is_stable_particle_index, gravity_field, img_shape, cur_config):
    """utils.py:336-356: one candidate per step x step window of the inner region."""
    by, bx = cur_config['boundary_size']['y'], cur_config['boundary_size']['x']
    step = cur_config['particle_call_window']
    thr = cur_config['particle_call_threshold']
    new = []
    for pix_y in range(by, img_shape[0] - by, step):
        for pix_x in range(bx, img_shape[1] - bx, step):
            win = gravity_field[pix_y:pix_y + step, pix_x:pix_x + step]
            a = np.unravel_index(win.argmin(), win.shape)
            if abs(win[a]) < thr:
                new.append((pix_x + a[1], pix_y + a[0]))
    if new:
        particles = np.vstack((particles, np.array(new, dtype=np.float64)))
    is_stable_particle_index = list(is_stable_particle_index) + [0] * len(new)
    return particles, is_stable_particle_index


def calc_distances(particles, GT_particles, n_neighbours, tie_mode='sklearn'):
    """utils.py:26-40: kNN(k+1), drop column 0, vec[i, j] = p_i - p_nbr(i, j)."""
    _, neigh_ind = knn(GT_particles, particles, n_neighbours + 1, tie_mode)
    neigh_ind = neigh_ind[:, 1:]
    top_dist_vectors = GT_particles[:, None, :] - particles[neigh_ind]
    return neigh_ind, top_dist_vectors


def calc_dist_energy_pit(particles, n_neighbours, mu, sigma, lambda_, use_optimisation=False,
                         is_stable_particle_index=None, tie_mode='sklearn'):
    """utils.py:112-129: F_i = sum_j vec/||vec|| * f(||vec||) over the k nearest."""
    particles = np.asarray(particles, dtype=np.float64)
    _, vec = calc_distances(particles, particles, n_neighbours, tie_mode)
    with np.errstate(invalid='ignore', divide='ignore'):
        # np.linalg.norm(vector) == sqrt(dot(v, v)); energy_pit_func2: sqrt(x**2. + y**2.)
        nrm = np.sqrt(vec[:, :, 0] * vec[:, :, 0] + vec[:, :, 1] * vec[:, :, 1])
        f = energy_pit_func2_closed(nrm, mu, sigma, lambda_)
        directions = (vec / nrm[:, :, None]) * f[:, :, None]
    out = np.zeros((particles.shape[0], 2))
    for j in range(n_neighbours):                      # np.sum(axis=0): row-sequential adds
        out = out + directions[:, j, :]
    if use_optimisation:
        out[np.asarray(is_stable_particle_index, dtype=bool)] = 0.
    with np.errstate(invalid='ignore'):
        modules = np.sqrt(out[:, 0] * out[:, 0] + out[:, 1] * out[:, 1])
    return out, modules


def calc_dist_energy_pit_optim(particles, is_stable_particle_index, n_neighbours, mu, sigma, lambda_,
                               tie_mode='sklearn'):
    """utils.py:132-134."""
    return calc_dist_energy_pit(particles, n_neighbours, mu, sigma, lambda_, use_optimisation=True,
                                is_stable_particle_index=is_stable_particle_index, tie_mode=tie_mode)


def calc_dist_energy(particles, n_neighbours, alpha, sigma, tie_mode='sklearn'):
    """utils.py:43-58 ('exp' energy): alpha * sum_j vec / (||vec|| / sigma^2)."""
    particles = np.asarray(particles, dtype=np.float64)
    _, vec = calc_distances(particles, particles, n_neighbours, tie_mode)
    with np.errstate(invalid='ignore', divide='ignore'):
        nrm = np.sqrt((vec[:, :, 0] ** 2.) + (vec[:, :, 1] ** 2.)) / (sigma ** 2.)
        directions = vec / nrm[:, :, None]
    out = np.zeros((particles.shape[0], 2))
    for j in range(n_neighbours):
        out = out + directions[:, j, :]
    out = alpha * out
    return out, np.sqrt(out[:, 0] * out[:, 0] + out[:, 1] * out[:, 1])


def move_particles(particles, is_stable_particle_index, blobness_grad, force, cur_config):
    """pipeline_with_delitions.py:215-245 (in place, Jacobi: forces precomputed)."""
    stable = np.asarray(is_stable_particle_index, dtype=bool)
    act = np.nonzero(~stable)[0]
    if act.size == 0:
        return particles
    p = particles
    g = blobness_grad[_trunc_index(p[act, 1]), _trunc_index(p[act, 0])]
    lb, ld = cur_config['lambda_blob'], cur_config['lambda_dist']
    if cur_config['step_mode'] == 'contin':
        dy = -1 * lb * g[:, 0] + -ld * force[act, 0]
        dx = -1 * lb * g[:, 1] + ld * force[act, 1]
        dy = -dy
        p[act, 0] += dx
        p[act, 1] += dy
    elif cur_config['step_mode'] == 'discrete':
        delta_y = -1. * lb * g[:, 0] + -ld * force[act, 1]
        delta_x = -1. * lb * g[:, 1] + ld * force[act, 0]
        delta_y = -delta_y
        ax, ay = np.abs(delta_x), np.abs(delta_y)
        with np.errstate(invalid='ignore'):
            sx, sy = np.sign(delta_x), np.sign(delta_y)
            xmaj = ax > ay
            step_x = np.where(xmaj, sx, np.where(ay < ax * 2, sx, 0.))
            step_y = np.where(xmaj, np.where(ax < ay * 2, sy, 0.), sy)
        p[act, 0] += step_x
        p[act, 1] += step_y
    return p


# --------------------------------------------------------------------------- L4 metrics
def get_particles_in_image(particles, cur_config, img_shape):
    """utils.py:404-418: strictly inside (b, H-b) x (b, W-b)."""
    by, bx = cur_config['boundary_size']['y'], cur_config['boundary_size']['x']
    x, y = particles[:, 0], particles[:, 1]
    keep = (bx < x) & (x < img_shape[1] - bx) & (by < y) & (y < img_shape[0] - by)
    idx = np.nonzero(keep)[0]
    return particles[idx], idx


def calc_acc_metrics(particles, GT_particles, gt_to_particles, particles_to_gt, cur_config, img_shape):
    """utils.py:378-402."""
    out = {}
    _, idx = get_particles_in_image(particles, cur_config, img_shape)
    for name, thr in (('1', 1.42), ('2', 2.83)):
        TP = np.sum([particles_to_gt <= thr])
        FN = np.sum([particles_to_gt > thr])
        FP = np.sum([gt_to_particles[idx] > thr])
        out[name] = {'TP': TP, 'FP': FP, 'FN': FN,
                     'dice': 2. * TP / (2. * TP + FP + FN),
                     'precision': TP / (TP + FP + 0.),
                     'recall': TP / (TP + FN + 0.)}
    return out


def calc_metrics(particles, GT_particles, gt_to_particles, particles_to_gt, mode='hangarian'):
    """utils.py:225-242."""
    if mode == 'hangarian':
        dist_matr = distance_matrix(GT_particles, particles)
        row_ind, col_ind = linear_sum_assignment(dist_matr)
        sel = dist_matr[row_ind, col_ind]
        return sel, sel.sum(), np.mean(sel), np.var(sel)
    elif mode == 'Chamfer':
        return (gt_to_particles, particles_to_gt), \
            np.sum(gt_to_particles) + np.sum(particles_to_gt), \
            np.mean(gt_to_particles) + np.mean(particles_to_gt), \
            np.var(gt_to_particles) + np.var(particles_to_gt)
    raise ValueError


def calc_blob_energy(R_b, particles):
    """utils.py:137-144."""
    return R_b[_trunc_index(particles[:, 1]), _trunc_index(particles[:, 0])]


# --------------------------------------------------------------------------- L5 orchestration
class Trace:
    def __init__(self):
        self.events = []

    def add(self, kind, **payload):
        self.events.append((kind, payload))

    def of(self, kind):
        return [p for k, p in self.events if k == kind]


def find_blobs(cur_config, GT_data, img, img_colorful=None, VERBOSE=False, USE_WANDB=False,
               experiment_idx=0, tie_mode='sklearn', trace=None, early_stop=True, scales=(1.0,)):
    """pipeline_with_delitions.py:12-403 (non-wandb, non-verbose branch)."""
    result = {'best_lsa_mean': +np.inf, 'best_blobEn_mean': -np.inf, 'dice_coeff_1': -np.inf,
              'dice_coeff_2': -np.inf, 'cur_config': cur_config, 'time_spended': 0}
    t0 = time.time()
    GT_data = GT_enumerate_from_zero(GT_data)
    GT_data = np.round(GT_data)
    img, img_colorful, GT_data = crop_image(img, img_colorful, GT_data, cur_config['region'])
    img, GT_data = mirror_padding_image(img, cur_config['boundary_size'], GT_data)

    tfield = time.time()
    R_b = calc_blobness(img, cur_config['blobness_formula'], scales)
    lut = get_precomputed_dist_func(cur_config)
    blobness_grad = calc_grad_field(R_b, cur_config['gradient_type'])
    field_time = time.time() - tfield
    if trace is not None:                                  # reference call order: lut (:41), grad (:91)
        trace.add('lut', lut=lut)
        trace.add('grad', R_b=R_b, grad=blobness_grad)

    particles = initialize_high_prop_location(R_b, cur_config['init_blob_threshold'], img.shape, cur_config)
    result['n_seeds'] = int(particles.shape[0])            # oracle-only diagnostics
    if trace is not None:
        trace.add('seeds', particles=np.array(particles))
    stable = [0] * particles.shape[0]

    def _delete(particles, stable, thr):
        pin = np.array(particles, dtype=np.float64)
        sin = np.array(stable, dtype=np.int8)
        new_p, deleted, new_s = del_close_particles(particles, stable, R_b, cur_config, threshold_dist=thr,
                                                    tie_mode=tie_mode)
        if trace is not None:
            trace.add('delete', particles_in=pin, stable_in=sin, threshold=float(thr),
                      particles_out=np.array(new_p), deleted=deleted, stable_out=np.array(new_s, dtype=np.int8))
        return new_p, deleted, new_s

    while True:
        particles, deleted, stable = _delete(particles, stable, 3)
        if deleted.sum() == 0:
            break
    stable = [1] * particles.shape[0]

    def _resample(particles, stable):
        field = calc_gravity_field_optim(lut, img.shape, particles, cur_config)
        if trace is not None:
            trace.add('gravity', particles=np.array(particles, dtype=np.float64), field=field)
        particles, stable = adding_particles(particles, stable, field, img.shape, cur_config)
        if trace is not None:
            trace.add('adding', particles_out=np.array(particles, dtype=np.float64),
                      stable_out=np.array(stable, dtype=np.int8))
        return particles, stable

    particles, stable = _resample(particles, stable)

    time_spended = {'dist_energy_calc': 0., 'moving_particles': 0., 'deleting_particles': 0.,
                    'fix_resample': 0., 'metrics_logging': 0.}
    n_iter = 0
    for iteration in range(cur_config['epoch_num']):
        n_iter += 1
        t1 = time.time()
        if cur_config['dist_energy_type'] == 'exp':
            force, modules = calc_dist_energy(particles, cur_config['dist_n_neighbours'],
                                              cur_config['dist_alpha'], cur_config['dist_sigma'], tie_mode)
            if trace is not None:
                trace.add('force_exp', particles=np.array(particles), force=force, modules=modules)
        elif cur_config['dist_energy_type'] == 'pit':
            # oracle-only diagnostics: energy_pit_func2 calls the reference's Python loop makes here (utils.py:117-127)
            result['n_energy_calls'] = result.get('n_energy_calls', int(cur_config['reception_field_size']) ** 2) + \
                (len(stable) - int(np.sum(stable))) * int(cur_config['dist_n_neighbours'])
            force, modules = calc_dist_energy_pit_optim(
                particles, stable, cur_config['dist_n_neighbours'], cur_config['mu_dist_func'],
                cur_config['sigma_dist_func'], cur_config['lambda_dist_func'], tie_mode)
            if trace is not None:
                trace.add('force', particles=np.array(particles), stable=np.array(stable, dtype=np.int8),
                          force=force, modules=modules)
        else:
            raise AttributeError('dist_energy_type={} is not implemented!'.format(cur_config['dist_energy_type']))
        time_spended['dist_energy_calc'] += time.time() - t1
        t1 = time.time()
        if cur_config['step_mode'] not in ('contin', 'discrete'):
            pass    # the reference silently does nothing for an unknown step_mode
        else:
            particles = move_particles(particles, stable, blobness_grad, force, cur_config)
        time_spended['moving_particles'] += time.time() - t1
        t1 = time.time()
        particles, _, stable = _delete(particles, stable, cur_config['threshhold_dist_del'])
        assert particles.shape[0] == len(stable)
        time_spended['deleting_particles'] += time.time() - t1
        t1 = time.time()
        if iteration % cur_config['fix_particles_frequency'] == 3:
            stable = [1] * len(stable)
            particles, stable = _resample(particles, stable)
        time_spended['fix_resample'] += time.time() - t1
        t1 = time.time()
        if iteration % cur_config['metric_measure_freq'] == 0:
            gt_to_particles = calc_dist_to_neares(GT_data, particles, 1, False, tie_mode)
            particles_to_gt = calc_dist_to_neares(particles, GT_data, 1, False, tie_mode)
            blob_energy = calc_blob_energy(R_b, particles)
            pm = calc_acc_metrics(particles, GT_data, gt_to_particles, particles_to_gt, cur_config, img.shape)
            if result['dice_coeff_1'] < pm['1']['dice'] * 0.995:
                result['dice_coeff_1'] = pm['1']['dice']
            if result['dice_coeff_2'] < pm['2']['dice'] * 0.995:
                result['dice_coeff_2'] = pm['2']['dice']
            lsa, lsa_sum, lsa_mean, lsa_var = calc_metrics(particles, GT_data, gt_to_particles, particles_to_gt,
                                                           mode=cur_config['metric_algo'])
            if result['best_lsa_mean'] > lsa_mean:
                result['best_lsa_mean'] = lsa_mean
            if result['best_blobEn_mean'] < np.mean(blob_energy):
                result['best_blobEn_mean'] = np.mean(blob_energy)
            result['last_metrics'] = {t: dict(v) for t, v in pm.items()}      # oracle-only diagnostics
            if trace is not None:                          # reference call order (:325, :329, :349)
                trace.add('blob_energy', energy=np.array(blob_energy))
                trace.add('acc', particles=np.array(particles), gt=np.array(GT_data),
                          gt_to_particles=gt_to_particles, particles_to_gt=particles_to_gt,
                          metrics={t: dict(v) for t, v in pm.items()})
                trace.add('chamfer', lsa_sum=float(lsa_sum), lsa_mean=float(lsa_mean), lsa_var=float(lsa_var))
        if early_stop and sum(stable) / len(particles) > 0.997:
            break
        time_spended['metrics_logging'] += time.time() - t1

    result['time_spended'] = time.time() - t0
    result['n_iterations'] = n_iter                   # oracle-only diagnostics
    result['final_particles'] = np.array(particles)
    result['final_stable'] = np.array(stable, dtype=np.int8)
    result['field_time'] = field_time
    return result, time_spended
