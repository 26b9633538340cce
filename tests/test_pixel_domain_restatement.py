"""CPU checks of the two reformulations the CUDA path rests on since round 2 (the kernels themselves are compared with
the oracle in test_gpu_parity.py; these tests pin the ARGUMENT, on the CPU, against the oracle):

  * gravity field: when the LUT is a function of dx^2 + dy^2 and non-increasing in it, the maximum of the stamped LUT
    entries over a pixel's window (reference utils.py:310-333) is the entry of the smallest squared distance among the
    particles in the window -- computed, like gravity_nearest_kernel, by a row pass on the window bits of a bitmap
    (two bit scans) and a min-plus column pass;
  * thinning: the loop `while deleted: del_close_particles(thr)` (reference :115-120, utils.py:245-278) on seeds in
    raster order is the same walk over pixels of a bitmap -- nearest alive neighbour = first alive pixel of the
    stencil sorted by (d^2, dy, dx), greedy order = raster order -- and its parallel resolution by wavefronts with
    (wave, ~pixel) stamps in two alternating buffers and one barrier per wave (thin_pixels_kernel), in any order inside
    a wave, gives the sequential result.
"""
import numpy as np

from oracle import aoslo_oracle as orc

M32 = 0xffffffff


def _ffs(v):
    return 0 if v == 0 else (v & -v).bit_length()


def _clz(v):
    return 32 - v.bit_length()


def _nearest_field(lut, shape, particles, rfs):
    H, W = shape
    h = rfs // 2
    T = {}
    for ky in range(rfs):
        for kx in range(rfs):
            d2 = (kx - h) ** 2 + (ky - h) ** 2
            assert T.setdefault(d2, lut[ky, kx]) == lut[ky, kx]          # radial
    ks = sorted(T)
    assert all(T[a] >= T[b] for a, b in zip(ks, ks[1:]))                 # non-increasing
    rows = [0] * H                                                        # one big-int bitmap row, bit x + 32
    for p in particles:
        rows[int(p[1])] |= 1 << (int(p[0]) + 32)
    INF = 0x4000
    field = np.empty(shape)
    r2 = np.empty((H, W), dtype=np.int64)
    for y in range(H):
        for x in range(W):
            w = (rows[y] >> (x + 32 - h)) & ((1 << (2 * h + 1)) - 1)      # window bits: pixels x-h .. x+h
            wr, wl = (w >> h), w & ((1 << h) - 1)
            rr = (_ffs(wr) - 1) & M32
            rl = (h - 31 + _clz(wl)) & M32
            r = min(rr, rl)
            r2[y, x] = r * r if r <= h else INF
    for y in range(H):
        for x in range(W):
            best = min(int(r2[yy, x]) + (yy - y) ** 2 for yy in range(max(0, y - h), min(H, y + h + 1)))
            field[y, x] = max(T[best], -0.48) if best < INF else -0.48
    return field


def test_windowed_distance_transform_is_the_gravity_field():
    rng = np.random.default_rng(0)
    for rfs, shape in ((21, (40, 90)), (19, (33, 70)), (5, (25, 40)), (31, (36, 64))):
        cfg = {'mu_dist_func': 5.84, 'sigma_dist_func': 0.83, 'lambda_dist_func': -0.1, 'reception_field_size': rfs}
        lut = orc.get_precomputed_dist_func(cfg)
        H, W = shape
        n = H * W // 30
        p = np.stack([rng.integers(0, W, n), rng.integers(0, H, n)], axis=1).astype(np.float64)
        p = np.vstack([p, [[0., 0.], [W - 1., H - 1.], [3.7, H - 2.2]]])
        np.testing.assert_array_equal(_nearest_field(lut, shape, p, rfs),
                                      orc.calc_gravity_field_optim(lut, shape, p, cfg))


def test_the_pit_lut_is_not_monotone_below_sigma_083():
    """why the general kernel exists: for sigma_dist_func < 0.83 (a third of the reference's sweep space) the pit
    function jumps upwards at mu and the LUT is not non-increasing in the distance"""
    cfg = {'mu_dist_func': 5.6, 'sigma_dist_func': 0.5, 'lambda_dist_func': -0.1, 'reception_field_size': 21}
    lut = orc.get_precomputed_dist_func(cfg)
    T = {}
    for ky in range(21):
        for kx in range(21):
            T[(kx - 10) ** 2 + (ky - 10) ** 2] = lut[ky, kx]
    ks = sorted(T)
    assert not all(T[a] >= T[b] for a, b in zip(ks, ks[1:]))


def _stencil():
    return [(dx, dy, d2) for d2 in range(1, 16) for dy in range(-3, 4) for dx in range(-3, 4)
            if dx * dx + dy * dy == d2]


def _thin_pixels(Rb, bx, by, seed_thr, thr, rng):
    H, W = Rb.shape
    st = _stencil()
    alive = np.zeros((H, W), bool)
    alive[by:H - by, bx:W - bx] = Rb[by:H - by, bx:W - bx] > seed_thr
    touch = [np.zeros((H, W), dtype=np.uint64), np.zeros((H, W), dtype=np.uint64)]
    wave = rounds = 0

    def stamp(buf, w, y, x, k):
        dx, dy, _ = st[k]
        key = np.uint64((w << 32) | (0xffffffff - (y * W + x)))
        buf[y, x] = max(buf[y, x], key)
        buf[y + dy, x + dx] = max(buf[y + dy, x + dx], key)

    while True:
        pend = []
        for y, x in zip(*np.nonzero(alive)):
            for k, (dx, dy, d2) in enumerate(st):
                if d2 >= thr * thr:
                    break
                if 0 <= x + dx < W and 0 <= y + dy < H and alive[y + dy, x + dx]:
                    pend.append((y, x, k))
                    break
        deleted = 0
        w_first = wave + 1
        for y, x, k in pend:
            stamp(touch[(wave + 1) & 1], wave + 1, y, x, k)
        prev = None
        while True:
            wave += 1
            if wave > w_first and prev == 0:
                break
            cur, nxt = touch[wave & 1], touch[(wave + 1) & 1]
            keep = []
            for i in rng.permutation(len(pend)):                          # any order inside a wave
                y, x, k = pend[i]
                dx, dy, _ = st[k]
                key = np.uint64((wave << 32) | (0xffffffff - (y * W + x)))
                if cur[y, x] == key and cur[y + dy, x + dx] == key:
                    if alive[y, x] and alive[y + dy, x + dx]:
                        if Rb[y, x] < Rb[y + dy, x + dx]:
                            alive[y, x] = False
                        else:
                            alive[y + dy, x + dx] = False
                        deleted += 1
                else:
                    stamp(nxt, wave + 1, y, x, k)
                    keep.append((y, x, k))
            pend, prev = keep, len(keep)
        rounds += 1
        if deleted == 0:
            break
    ys, xs = np.nonzero(alive)
    return np.stack([xs, ys], axis=1).astype(np.float64), rounds


def test_pixel_domain_thinning_is_the_reference_loop():
    rng = np.random.default_rng(3)
    for H, W, thr, seed_thr in ((48, 70, 3.0, 0.3), (40, 56, 3.0, 0.7), (36, 50, 3.5, 0.5), (36, 60, 2.0, 0.4),
                                (32, 44, 4.0, 0.6)):
        Rb = np.round(rng.random((H, W)) * 20) / 20                       # quantised: equal energies occur
        bx, by = 4, 3
        cfg = {'boundary_size': {'x': bx, 'y': by}, 'verbose_func': False}
        p = orc.initialize_high_prop_location(Rb, seed_thr, Rb.shape, cfg)
        stable, rounds = [0] * p.shape[0], 0
        while True:
            p, deleted, stable = orc.del_close_particles(p, stable, Rb, cfg, threshold_dist=thr, tie_mode='canonical')
            rounds += 1
            if int(np.sum(deleted)) == 0:
                break
        got, got_rounds = _thin_pixels(Rb, bx, by, seed_thr, thr, rng)
        np.testing.assert_array_equal(got, p)
        assert got_rounds == rounds
