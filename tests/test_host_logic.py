"""CPU tests: C-ABI surface, config packing, host geometry, sharding + gloo all-gather (world_size 2)."""
import os
import re
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from detecting_cones_aoslo_b200 import _lib
    header = open(os.path.join(ROOT, "include", "aoslo_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(aoslo_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))
    for name in declared:
        assert getattr(_lib.lib, name) is not None           # bound by ctypes at import
    assert _lib.lib.aoslo_version().decode().startswith("aoslo_b200")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "detecting_cones_aoslo_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from detecting_cones_aoslo_b200 import utils
    with pytest.raises(RuntimeError):
        utils.calc_blobness(np.zeros((16, 16), np.uint8) + 7, {'blobness_formula': 'custom',
                                                               'gradient_type': 'np_grad'})


def test_crop_and_pad_match_oracle():
    from detecting_cones_aoslo_b200 import image_transformation as it
    from oracle import aoslo_oracle as orc
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(60, 70), dtype=np.uint8)
    gt = rng.random((200, 2)) * [70, 60]
    region = {'x_min': 5, 'x_max': 61, 'y_min': 3, 'y_max': 50}
    a_img, _, a_gt = it.crop_image(img, None, gt.copy(), region)
    b_img, _, b_gt = orc.crop_image(img, None, gt.copy(), region)
    np.testing.assert_array_equal(a_img, b_img)
    np.testing.assert_array_equal(a_gt, b_gt)
    pa, ga = it.mirror_padding_image(a_img, {'x': 7, 'y': 4}, a_gt.copy())
    pb, gb = orc.mirror_padding_image(b_img, {'x': 7, 'y': 4}, b_gt.copy())
    np.testing.assert_array_equal(pa, pb)
    np.testing.assert_array_equal(ga, gb)
    with pytest.raises(AssertionError):
        it.crop_image(img, None, gt.copy(), {'x_min': 0, 'x_max': 70, 'y_min': 0, 'y_max': 10})


def test_config_struct_packing_and_errors():
    from detecting_cones_aoslo_b200 import configs
    from detecting_cones_aoslo_b200.pipeline_with_delitions import make_config_struct
    c = make_config_struct(configs.zoo_config(), 549, 552)
    assert (c.H, c.W, c.bx, c.by) == (549, 552, 17, 17)
    assert c.dist_n_neighbours == 3 and c.reception_field_size == 21 and c.tie_mode == 1 and c.early_stop == 1
    assert c.threshhold_dist_del == 3.5 and c.thinning_threshold == 3.0
    with pytest.raises(TypeError):                      # float window: range() TypeError in the reference
        make_config_struct(configs.variant(configs.zoo_config(), particle_call_window=5.5), 549, 552)
    with pytest.raises(AttributeError):
        make_config_struct(configs.variant(configs.zoo_config(), dist_energy_type='nope'), 549, 552)


def test_sweep_sampler_is_deterministic_and_in_range():
    from detecting_cones_aoslo_b200 import configs, sweep
    a = sweep.sample_trial(configs.zoo_config(), 17)
    b = sweep.sample_trial(configs.zoo_config(), 17)
    assert a == b
    for i in range(50):
        c = sweep.sample_trial(configs.zoo_config(), i)
        assert 0.05 <= c['lambda_dist'] <= 2.0 and 1 <= c['dist_n_neighbours'] <= 6
        assert isinstance(c['particle_call_window'], int) and 5 <= c['particle_call_window'] <= 9
        assert 4 <= c['fix_particles_frequency'] <= 8 and c['reception_field_size'] == 19


def test_shards_partition_the_units():
    from detecting_cones_aoslo_b200 import sweep
    for n, w in ((1024, 8), (256, 8), (10, 4), (3, 8), (7, 2)):
        seen = [u for r in range(w) for u in sweep.shard_units(n, r, w)]
        assert seen == list(range(n))


def _worker(rank, world, port, n_units, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from detecting_cones_aoslo_b200 import sweep
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def run_unit(u):
        if u == 5:
            raise RuntimeError("a failing trial must not kill the rank")
        rng = np.random.default_rng(u)
        rec = np.zeros(sweep.RECORD_LEN)
        rec[0] = u
        rec[1:] = rng.integers(0, 10000, sweep.RECORD_LEN - 1)
        return rec

    table = sweep.run_sharded(sweep.shard_units(n_units, rank, world), run_unit, n_units)
    np.save(os.path.join(out_dir, "table_%d.npy" % rank), table)
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_rank(tmp_path):
    """N>1 path on CPU: gloo, world_size 2; the gathered table equals the 1-rank table bit for bit."""
    import torch.multiprocessing as mp
    from detecting_cones_aoslo_b200 import sweep
    n_units = 11
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n_units, str(tmp_path)), nprocs=2, join=True)
    t0 = np.load(tmp_path / "table_0.npy")
    t1 = np.load(tmp_path / "table_1.npy")

    def run_unit(u):
        if u == 5:
            raise RuntimeError
        rng = np.random.default_rng(u)
        rec = np.zeros(sweep.RECORD_LEN)
        rec[0] = u
        rec[1:] = rng.integers(0, 10000, sweep.RECORD_LEN - 1)
        return rec

    single = sweep.run_sharded(range(n_units), run_unit, n_units)
    assert t0.shape == (n_units, sweep.RECORD_LEN)
    np.testing.assert_array_equal(t0, t1)
    np.testing.assert_array_equal(t0, single)
    assert t0[5, 1] == -1 and np.isnan(t0[5, 2])           # the failed unit is recorded, not fatal


def test_u8_over_255_by_reciprocal_and_one_fma_correction_is_correctly_rounded():
    """blobness_dtile_kernel evaluates img_as_float's v / 255 as q = v * r, q + fma(-q, 255, v) * r with
    r = fl(1 / 255); exact rational arithmetic shows that this is the correctly rounded quotient for every byte
    (the plain product v * r is wrong for 24 of them)."""
    from fractions import Fraction
    r = 1.0 / 255.0
    wrong_plain = 0
    for v in range(256):
        exact = float(Fraction(v, 255))                       # float(Fraction) rounds to nearest even
        q = float(v) * r
        wrong_plain += q != exact
        e = float(Fraction(v) - Fraction(q) * 255)            # fma(-q, 255, v): one rounding
        q2 = float(Fraction(e) * Fraction(r) + Fraction(q))   # fma(e, r, q): one rounding
        assert q2 == exact == v / 255.0
    assert wrong_plain == 24


def test_lsap_restatement_equals_scipy_on_tie_heavy_instances():
    """csrc/hungarian.cu restates scipy's linear_sum_assignment (an un-vendored dependency of the reference's
    calc_metrics('hangarian'), utils.py:228-234); the same restatement in NumPy returns scipy's assignment --
    not just its cost -- on integer matrices and lattice distances, where optimal assignments are far from unique."""
    from scipy.optimize import linear_sum_assignment
    from lsap_restated import lsa_restated
    rng = np.random.default_rng(0)
    for t in range(240):
        nr, nc = rng.integers(1, 36, 2)
        if t % 3 == 0:
            c = rng.integers(0, 4, (nr, nc)).astype(float)
        elif t % 3 == 1:
            a, b = rng.integers(0, 12, (nr, 2)), rng.integers(0, 12, (nc, 2))
            c = np.sqrt(((a[:, None, :] - b[None, :, :]) ** 2).sum(-1).astype(float))
        else:
            c = rng.random((nr, nc))
        r0, c0 = linear_sum_assignment(c)
        r1, c1, _ = lsa_restated(c)
        assert np.array_equal(r0, r1) and np.array_equal(c0, c1), (t, nr, nc)


# ----------------------------------------------------------------------------------- input path (f2)
def _packbits(row):
    """TIFF PackBits encoder (test-side; the product only decodes)."""
    out, i, n = bytearray(), 0, len(row)
    while i < n:
        j = i
        while j + 1 < n and row[j + 1] == row[i] and j - i < 127:
            j += 1
        run = j - i + 1
        if run >= 3:
            out += bytes([(257 - run) & 0xff, row[i]])
            i = j + 1
            continue
        k = i
        while k < n and k - i < 128:
            if k + 2 < n and row[k] == row[k + 1] == row[k + 2]:
                break
            k += 1
        out += bytes([k - i - 1]) + bytes(row[i:k])
        i = k
    return bytes(out)


def _write_tiff(path, img, compression=1, big_endian=False, rows_per_strip=15):
    import struct
    bo = '>' if big_endian else '<'
    H, W = img.shape
    strips = []
    for r0 in range(0, H, rows_per_strip):
        rows = img[r0:r0 + rows_per_strip]
        strips.append(rows.tobytes() if compression == 1 else b''.join(_packbits(bytes(r)) for r in rows))
    data = b''.join(strips)
    n = len(strips)
    head = (b'MM' if big_endian else b'II') + struct.pack(bo + 'HI', 42, 8)
    tags = [(256, 3, 1, W), (257, 3, 1, H), (258, 3, 1, 8), (259, 3, 1, compression), (262, 3, 1, 1), (273, 4, n, None),
            (277, 3, 1, 1), (278, 3, 1, rows_per_strip), (279, 4, n, None)]
    ifd_size = 2 + 12 * len(tags) + 4
    off_arrays = 8 + ifd_size
    off_data = off_arrays + 8 * n
    offs, o = [], off_data
    for s in strips:
        offs.append(o)
        o += len(s)
    ifd = struct.pack(bo + 'H', len(tags))
    for tag, typ, cnt, val in tags:
        if val is None:
            arr_off = off_arrays + (0 if tag == 273 else 4 * n)
            if cnt == 1:
                val = offs[0] if tag == 273 else len(strips[0])
            else:
                val = arr_off
            ifd += struct.pack(bo + 'HHII', tag, typ, cnt, val)
        elif typ == 3:
            ifd += struct.pack(bo + 'HHIHH', tag, typ, cnt, val, 0)
        else:
            ifd += struct.pack(bo + 'HHII', tag, typ, cnt, val)
    ifd += struct.pack(bo + 'I', 0)
    arrays = struct.pack(bo + '%dI' % n, *offs) + struct.pack(bo + '%dI' % n, *[len(s) for s in strips])
    with open(path, 'wb') as f:
        f.write(head + ifd + arrays + data)


@pytest.mark.parametrize("compression,big_endian,rps", [(1, False, 15), (32773, False, 15), (32773, True, 7), (1, True, 600)])
def test_tiff_and_csv_readers(tmp_path, compression, big_endian, rps):
    """The native readers (csrc/io_host.cpp) on the shipped image re-encoded by this test: uncompressed / PackBits
    (what the shipped TIFF uses), both byte orders, one or many strips; CSV with and without the header-row quirk."""
    from golden_io import load_dataset
    from detecting_cones_aoslo_b200 import dataio
    img, gt = load_dataset()
    tp, cp = str(tmp_path / "a.tiff"), str(tmp_path / "a.csv")
    _write_tiff(tp, img, compression, big_endian, rps)
    assert dataio.tiff_shape(tp)[:2] == img.shape
    np.testing.assert_array_equal(dataio.read_tiff(tp), img)
    full = np.vstack([[1.5, 30.0], gt])                      # the first line pd.read_csv eats as a header
    with open(cp, 'w') as f:
        for x, y in full:
            f.write("%r,%r\r\n" % (float(x), float(y)))
    np.testing.assert_array_equal(dataio.read_points_csv(cp), gt)                       # quirk on: line 1 dropped
    np.testing.assert_array_equal(dataio.read_points_csv(cp, header_row_quirk=False), full)
    with pytest.raises(ValueError):
        dataio.read_tiff(str(tmp_path / "missing.tiff"))
    bad = str(tmp_path / "bad.csv")
    open(bad, 'w').write("1,2\nx,y\n")
    with pytest.raises(ValueError):
        dataio.read_points_csv(bad, header_row_quirk=False)


def test_readers_equal_cv2_and_pandas_on_the_shipped_files():
    """In the build container the reference tree is present: the native readers must return exactly what the
    reference's own loaders (cv2.imread(path, -1), pd.read_csv(path).to_numpy()) return, = tests/golden/dataset.npz."""
    import glob
    tiffs = glob.glob("/root/reference/dataset/*.tiff")
    if not tiffs:
        pytest.skip("reference dataset not present (GPU box)")
    from golden_io import load_dataset
    from detecting_cones_aoslo_b200 import dataio
    img, gt = load_dataset()
    i2, g2 = dataio.load_pair(tiffs[0], tiffs[0][:-5] + ".csv")
    np.testing.assert_array_equal(i2, img)
    np.testing.assert_array_equal(g2, gt)
    assert dataio.read_points_csv(tiffs[0][:-5] + ".csv", header_row_quirk=False).shape[0] == gt.shape[0] + 1


# ----------------------------------------------------------------------------------- sweep space (f1)
SWEEP_YAML = """
program: main_with_delitions.py
method: bayes
metric: {name: 2_dice, goal: maximize}
parameters:
  region.x_max: {value: 518}
  region.x_min: {value: 0}
  region.y_max: {value: 515}
  region.y_min: {value: 0}
  boundary_size.x: {value: 15}
  boundary_size.y: {value: 15}
  epoch_num: {value: 30}
  write_gif: {value: 0}
  step_mode: {distribution: categorical, values: [discrete]}
  lambda_dist: {distribution: uniform, min: 0.05, max: 2.}
  dist_n_neighbours: {distribution: int_uniform, min: 1, max: 6}
  particle_call_window: {distribution: categorical, values: [5., 5.5, 6., 6.4, 9.]}
"""


def test_sweep_yaml_space_and_grid_driver(tmp_path):
    from detecting_cones_aoslo_b200 import configs, sweep
    y = tmp_path / "sweep.yaml"
    y.write_text(SWEEP_YAML)
    space, fixed = sweep.load_sweep_space(str(y))
    assert fixed['region'] == {'x_max': 518, 'x_min': 0, 'y_max': 515, 'y_min': 0}
    assert fixed['boundary_size'] == {'x': 15, 'y': 15} and fixed['step_mode'] == 'discrete'
    assert fixed['write_gif'] is False and fixed['epoch_num'] == 30
    assert space['lambda_dist'] == ('uniform', 0.05, 2.0) and space['dist_n_neighbours'] == ('int_uniform', 1, 6)
    assert space['particle_call_window'] == ('choice', (5, 6, 9))      # the float members crash the reference
    space_f, _ = sweep.load_sweep_space(str(y), integer_windows=False)
    assert 5.5 in space_f['particle_call_window'][1]
    a = sweep.sample_trial(configs.zoo_config(), 5, space=space, fixed=fixed)
    b = sweep.sample_trial(configs.zoo_config(), 5, space=space, fixed=fixed)
    assert a == b and a['boundary_size'] == {'x': 15, 'y': 15} and isinstance(a['particle_call_window'], int)
    assert 0.05 <= a['lambda_dist'] <= 2.0 and 1 <= a['dist_n_neighbours'] <= 6
    # the built-in space restates the reference's file: compare when the reference tree is present
    ref = "/root/reference/sweep_config_new.yaml"
    if os.path.exists(ref):
        rs, rf = sweep.load_sweep_space(ref)
        assert rs == sweep.SWEEP_SPACE
        for k, v in sweep.SWEEP_FIXED.items():
            if k in rf:
                assert rf[k] == v, k
    # grid driver helpers (reference main_with_delitions.py:17-26)
    zoo = {'a': [1, 2, 3], 'b': 'x', 'c': [0.1, 0.2]}
    combos = list(sweep.grid_parameters(zoo))
    assert len(combos) == sweep.estimate_num_variant(zoo) == 6 and combos[0] == {'a': 1, 'b': 'x', 'c': 0.1}
    # records -> benchmarking_chamfer.csv
    table = np.array([sweep.pack_record(i, {'dice_coeff_1': 0.5 + 0.1 * i, 'dice_coeff_2': 0.9, 'best_lsa_mean': 1.0,
                                            'best_blobEn_mean': 5.0}, seconds=0.01) for i in (2, 0, 1)])
    res = sweep.table_to_results(table, {i: {'k': i} for i in range(3)})
    out = sweep.write_results_table(res, str(tmp_path / "benchmarking_chamfer.csv"))
    assert list(out['dice_coeff_1']) == sorted(out['dice_coeff_1'])
    assert set(out.columns) == {'best_lsa_mean', 'best_blobEn_mean', 'dice_coeff_1', 'dice_coeff_2', 'cur_config',
                                'time_spended'}
