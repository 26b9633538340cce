"""ctypes binding of libaoslo_b200.so (include/aoslo_b200.h).

There is NO fallback: if the shared library is missing or a symbol cannot be
bound, importing this module raises.  Build it with
`python -m detecting_cones_aoslo_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libaoslo_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_UNSUPPORTED = 0, -1, -2, -3, -4
DEV_ESCAPED, DEV_NAN, DEV_CAPACITY, DEV_EIG_ASSERT, DEV_KNN_SHORT = 1, 2, 4, 8, 16
F32, F64 = 0, 1
FORMULA = {'custom': 0, 'simple_div': 1}
GRAD = {'np_grad': 0, 'sobel': 1}
STEP = {'discrete': 0, 'contin': 1}
ENERGY = {'pit': 0, 'exp': 1}
TIE = {'canonical': 0, 'sklearn': 1}
KMAX = 8
MAX_SCALES = 8
MAX_RADIUS = 16


class Config(C.Structure):
    _fields_ = [
        ("H", C.c_int32), ("W", C.c_int32), ("bx", C.c_int32), ("by", C.c_int32),
        ("lambda_dist", C.c_double), ("lambda_blob", C.c_double),
        ("dist_alpha", C.c_double), ("dist_sigma", C.c_double),
        ("dist_n_neighbours", C.c_int32), ("epoch_num", C.c_int32),
        ("metric_measure_freq", C.c_int32), ("step_mode", C.c_int32), ("dist_energy_type", C.c_int32),
        ("mu_dist_func", C.c_double), ("sigma_dist_func", C.c_double), ("lambda_dist_func", C.c_double),
        ("threshhold_dist_del", C.c_double),
        ("particle_call_window", C.c_int32),
        ("particle_call_threshold", C.c_double), ("init_blob_threshold", C.c_double),
        ("reception_field_size", C.c_int32), ("fix_particles_frequency", C.c_int32),
        ("tie_mode", C.c_int32), ("early_stop", C.c_int32),
        ("thinning_threshold", C.c_double),
    ]


class MetricsRecord(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("n_particles", C.c_int32), ("n_stable", C.c_int32), ("n_gt", C.c_int32),
        ("tp", C.c_int32 * 2), ("fp", C.c_int32 * 2), ("fn", C.c_int32 * 2),
        ("n_in_image", C.c_int32), ("pad_", C.c_int32),
        ("sum_p2g", C.c_double), ("sum_g2p", C.c_double),
        ("var_p2g", C.c_double), ("var_g2p", C.c_double),
        ("blob_energy_sum", C.c_double),
    ]


class RunSummary(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("n_iterations", C.c_int32), ("n_particles", C.c_int32), ("n_stable", C.c_int32),
        ("n_seeds", C.c_int32), ("n_after_thinning", C.c_int32), ("n_thinning_rounds", C.c_int32),
        ("n_records", C.c_int32),
    ]


class TrialResult(C.Structure):
    _fields_ = [
        ("error", C.c_int32), ("pad_", C.c_int32),
        ("summary", RunSummary), ("last", MetricsRecord),
        ("dice_coeff_1", C.c_double), ("dice_coeff_2", C.c_double),
        ("best_lsa_mean", C.c_double), ("best_blobEn_mean", C.c_double),
    ]


# every symbol include/aoslo_b200.h declares: name -> (restype, argtypes)
_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_ip = C.POINTER(C.c_int)
SYMBOLS = {
    "aoslo_version": (C.c_char_p, []),
    "aoslo_last_cuda_error": (C.c_char_p, []),
    "aoslo_device_count": (_i, []),
    "aoslo_launch_count": (C.c_longlong, []),
    "aoslo_set_wait_policy": (_i, [_i, _i]),
    "aoslo_ctx_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "aoslo_ctx_destroy": (_i, [_vp]),
    "aoslo_ctx_status": (_i, [_vp, _vp, _ip, _i]),
    "aoslo_ctx_clear_status": (_i, [_vp, _vp]),
    "aoslo_ctx_capacity": (_i, [_vp]),
    "aoslo_ctx_set_blocking_sync": (_i, [_vp, _i]),
    "aoslo_ctx_set_tuning": (_i, [_vp, C.c_char_p, _i]),
    "aoslo_prepare_inputs": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _i]),
    "aoslo_stage_gt": (_i, [_vp, _vp, _vp, _i, _vp, _i]),
    "aoslo_ctx_flush_host_work": (_i, [_vp]),
    "aoslo_blobness": (_i, [_vp, _vp, _vp, _i, _i, _i, C.POINTER(_d), _ip, C.POINTER(_d), _i, _i, _i, _i, _vp, _vp]),
    "aoslo_grad_field": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "aoslo_seed": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _i, _ip]),
    "aoslo_knn": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    "aoslo_delete_close": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _d, _i, _vp, _vp, _vp, _ip]),
    "aoslo_dist_lut": (_i, [_vp, _vp, _i, _d, _d, _d, _vp]),
    "aoslo_gravity_field": (_i, [_vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "aoslo_add_particles": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _vp, _i, _i, _ip]),
    "aoslo_dist_force": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _d, _d, _d, _i, _vp]),
    "aoslo_move": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp, _d, _d, _i]),
    "aoslo_metrics": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, C.POINTER(MetricsRecord),
                           _vp, _vp]),
    "aoslo_classify": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _d, _vp, _vp]),
    "aoslo_hungarian": (_i, [_vp, _vp, _vp, _i, _vp, _i, _vp, _vp]),
    "aoslo_iterate": (_i, [_vp, _vp, C.POINTER(Config), _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _ip]),
    "aoslo_run": (_i, [_vp, _vp, C.POINTER(Config), _vp, _vp, _i, _vp, _i, C.POINTER(MetricsRecord), _i, _vp, _vp,
                       C.POINTER(RunSummary), C.POINTER(C.c_float)]),
    "aoslo_sweep": (_i, [_vp, _vp, C.POINTER(Config), _i, _vp, _vp, _i, _vp, _i, C.POINTER(TrialResult)]),
    "aoslo_read_particles": (_i, [_vp, _vp, _vp, _vp, _i]),
    "aoslo_tiff_info": (_i, [C.c_char_p, _ip, _ip, _ip, _ip]),
    "aoslo_tiff_read_gray8": (_i, [C.c_char_p, _vp, _i]),
    "aoslo_csv_read_points": (_i, [C.c_char_p, _i, _vp, _i, _ip]),
    "aoslo_kdtree_build_host": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i]),
    "aoslo_kdtree_num_nodes": (_i, [_i]),
}


def load(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            "libaoslo_b200.so not found at %s -- the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m detecting_cones_aoslo_b200.build`." % path)
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load()


class AosloError(RuntimeError):
    pass


def check(status, what=""):
    if status == OK:
        return
    if status == ERR_CUDA:
        raise AosloError("%s: CUDA error: %s" % (what, lib.aoslo_last_cuda_error().decode()))
    if status == ERR_CAPACITY:
        raise MemoryError("%s: particle capacity of the context exceeded" % what)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError("%s: option not implemented by libaoslo_b200" % what)
    raise ValueError("%s: invalid argument (status %d)" % (what, status))


def raise_device_status(word):
    """Map the device status word to the exception the reference would raise."""
    if word & DEV_EIG_ASSERT:
        raise AssertionError("(eigs[0] > eigs[1]).all() violated (reference pipeline_with_delitions.py:43)")
    if word & DEV_NAN:
        raise ValueError("cannot convert float NaN to integer (coincident particles -> 0/0 force, "
                         "reference utils.py:123 / pipeline_with_delitions.py:243)")
    if word & DEV_ESCAPED:
        raise IndexError("particle left the padded image (reference indexes blobness_grad out of bounds, "
                         "pipeline_with_delitions.py:230)")
    if word & DEV_CAPACITY:
        raise MemoryError("particle capacity exceeded on the device")
    if word & DEV_KNN_SHORT:
        raise ValueError("Expected n_neighbors <= n_samples_fit (sklearn NearestNeighbors)")
