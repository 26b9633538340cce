"""Named detection configurations (plain dicts, the reference's config surface).

Key set = the dict literal of the reference's pipeline_with_delitions.py:411-436
(29 keys incl. the `threshhold_dist_del` spelling).  Values of the named presets
cite where the reference defines them.
"""
import copy

FULL_REGION = {'x_min': 0, 'x_max': 518, 'y_min': 0, 'y_max': 515}


def _base():
    return {
        'random_seed': 2022, 'lambda_dist': 1., 'lambda_blob': 0.,
        'dist_alpha': 0.3, 'dist_sigma': 0.3, 'dist_n_neighbours': 1,
        'region': {'x_min': 50, 'x_max': 150, 'y_min': 50, 'y_max': 150},
        'boundary_size': {'x': 10, 'y': 10},
        'gradient_type': 'np_grad',
        'epoch_num': 30, 'n_particles_coeff': 1.0, 'metric_measure_freq': 3,
        'step_mode': 'discrete',
        'blobness_formula': 'custom',
        'write_gif': False,
        'metric_algo': 'Chamfer',
        'mu_dist_func': 5.84,
        'sigma_dist_func': 0.83,
        'lambda_dist_func': -0.1,
        'dist_energy_type': 'pit',
        'threshhold_dist_del': 3.5,
        'particle_call_window': 5,
        'particle_call_threshold': 0.1,
        'init_blob_threshold': 5.,
        'reception_field_size': 21,
        'fix_particles_frequency': 4,
        'verbose_func': False,
        'save_best_result_visualisation': False,
    }


def main_config():
    """`__main__` of the reference's pipeline_with_delitions.py:411-441
    (100x100 crop, boundary 10, lambda_blob 0, k=1); viz flags off."""
    return _base()


def zoo_config():
    """`config_zoo` of the reference's main_with_delitions.py:35-75
    (full region, boundary 17, k=3, call threshold 0.3)."""
    c = _base()
    c.update({'lambda_dist': 0.1, 'lambda_blob': 1.0, 'dist_n_neighbours': 3,
              'region': dict(FULL_REGION), 'boundary_size': {'x': 17, 'y': 17},
              'particle_call_threshold': 0.3})
    return c


def full_b10_config():
    """Full region, boundary 10, k=3, lambda_blob 1, lambda_dist .1, call thr .1
    (golden row 3 of BASELINE.md section 2)."""
    c = _base()
    c.update({'lambda_dist': 0.1, 'lambda_blob': 1.0, 'dist_n_neighbours': 3,
              'region': dict(FULL_REGION), 'boundary_size': {'x': 10, 'y': 10}})
    return c


def sweep_like_config():
    """One point of the reference's sweep_config_new.yaml space
    (boundary 15, rfs 19, window 9, freq 6, thr 4.98, del 3.0, lambda_dist .5, k=3)."""
    c = _base()
    c.update({'lambda_dist': 0.5, 'lambda_blob': 1.0, 'dist_n_neighbours': 3,
              'region': dict(FULL_REGION), 'boundary_size': {'x': 15, 'y': 15},
              'reception_field_size': 19, 'particle_call_window': 9,
              'fix_particles_frequency': 6, 'init_blob_threshold': 4.98,
              'threshhold_dist_del': 3.0})
    return c


def variant(base, **over):
    c = copy.deepcopy(base)
    c.update(over)
    return c


NAMED = {
    'main': main_config,
    'zoo': zoo_config,
    'full_b10': full_b10_config,
    'sweep_like': sweep_like_config,
}
