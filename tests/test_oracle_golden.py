"""Pins the CPU oracle (oracle/aoslo_oracle.py) against traces of the UNMODIFIED
reference executed in the build container (tests/golden/, oracle/gen_golden.py)."""
import numpy as np
import pytest

from detecting_cones_aoslo_b200 import configs
from oracle import aoslo_oracle as orc
from golden_io import GoldenTrace, load_dataset
from trace_compare import compare_traces

CASES = {
    "main": configs.main_config,
    "main_sobel": lambda: configs.variant(configs.main_config(), gradient_type="sobel", lambda_blob=1.),
    "main_contin": lambda: configs.variant(configs.main_config(), step_mode="contin", lambda_blob=1.),
    "main_exp": lambda: configs.variant(configs.main_config(), dist_energy_type="exp"),
    "main_hangarian": lambda: configs.variant(configs.main_config(), metric_algo="hangarian"),
    "main_k3": lambda: configs.variant(configs.main_config(), dist_n_neighbours=3, lambda_blob=1., lambda_dist=0.1),
    "zoo": configs.zoo_config,
    "full_b10": configs.full_b10_config,
    "sweep_like": configs.sweep_like_config,
}


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_replays_reference_trace(name):
    img, gt = load_dataset()
    gold = GoldenTrace(name)
    tr = orc.Trace()
    result, _ = orc.find_blobs(CASES[name](), gt.copy(), img, None, False, False, tie_mode="sklearn", trace=tr)
    # fields: the oracle runs the same scipy/numpy calls as the reference -> exact;
    # forces: closed-form energy vs scipy.stats frozen dists, <= 1 ulp of O(1) values
    compare_traces(gold.events, tr.events, stride=gold.stride, field_tol=0.0, force_tol=2e-15, label=name)
    for key, val in gold.result.items():
        np.testing.assert_allclose(float(result[key]), val, rtol=1e-13, err_msg=key)


def test_reference_published_goldens():
    """Numbers quoted in BASELINE.md section 2 (reference executed, `__main__` config)."""
    gold = GoldenTrace("main")
    counts = [p["particles_in"].shape[0] for p in gold.of("delete") if float(p["threshold"]) == 3.5]
    assert counts == [283, 277, 275, 275]
    acc = gold.of("acc")
    m0, m3 = acc[0]["metrics"], acc[1]["metrics"]
    assert (m0["1"]["TP"], m0["1"]["FP"], m0["1"]["FN"]) == (229, 32, 24)
    assert (m0["2"]["TP"], m0["2"]["FP"], m0["2"]["FN"]) == (244, 14, 9)
    assert (m3["1"]["TP"], m3["1"]["FP"], m3["1"]["FN"]) == (234, 24, 19)
    assert (m3["2"]["TP"], m3["2"]["FP"], m3["2"]["FN"]) == (241, 15, 12)
    assert m0["2"]["dice"] == 0.9549902152641878 and m0["1"]["dice"] == 0.8910505836575876
    assert gold.result["best_lsa_mean"] == 1.9616404226188986
    assert gold.result["best_blobEn_mean"] == 5.030508937676305
    zoo = GoldenTrace("zoo")
    assert zoo.of("seeds")[0]["particles"].shape[0] == 31876
    assert [p["particles_out"].shape[0] for p in zoo.of("delete")][:5] == [13573, 6739, 5569, 5521, 5521]
    assert zoo.result["dice_coeff_2"] == 0.9343206696716033
    assert zoo.result["dice_coeff_1"] == 0.833386591678916
