"""Event-by-event comparison of two find_blobs traces (golden / oracle / GPU)."""
import numpy as np


def _sub(a, stride):
    a = np.asarray(a)
    return a[::stride, ::stride] if stride > 1 else a


def compare_traces(gold_events, got_events, stride=1, field_tol=0.0, force_tol=1e-12, label=""):
    """gold_events / got_events: lists of (kind, payload).  `stride` sub-samples the
    dense fields of `got` to match a strided golden.  Discrete quantities (particle
    arrays, masks, counts) must match exactly; FP fields to `field_tol` (abs)."""
    gk = [k for k, _ in gold_events]
    ok = [k for k, _ in got_events]
    assert gk == ok, "%s event kinds differ:\n gold=%s\n got =%s" % (label, gk, ok)
    for i, ((kind, g), (_, o)) in enumerate(zip(gold_events, got_events)):
        tag = "%s event %d (%s)" % (label, i, kind)
        if kind == "grad":
            for key in ("R_b", "grad"):
                a, b = np.asarray(g[key], dtype=np.float64), _sub(o[key], stride)
                assert a.shape == b.shape, tag
                err = np.abs(a - b).max()
                assert err <= field_tol, "%s %s err %g" % (tag, key, err)
        elif kind == "lut":
            np.testing.assert_allclose(o["lut"], g["lut"], rtol=0, atol=max(field_tol, 3e-16), err_msg=tag)
        elif kind == "seeds":
            np.testing.assert_array_equal(np.asarray(o["particles"]), np.asarray(g["particles"]), err_msg=tag)
        elif kind == "delete":
            assert float(g["threshold"]) == float(o["threshold"]), tag
            for key in ("particles_in", "particles_out", "stable_in", "stable_out", "deleted"):
                a, b = np.asarray(g[key]), np.asarray(o[key])
                assert a.shape == b.shape, "%s %s shape %s vs %s" % (tag, key, a.shape, b.shape)
                if a.dtype.kind == "f" or b.dtype.kind == "f":
                    np.testing.assert_allclose(b.astype(np.float64), a.astype(np.float64), rtol=0,
                                               atol=force_tol, err_msg="%s %s" % (tag, key))
                else:
                    np.testing.assert_array_equal(b, a, err_msg="%s %s" % (tag, key))
        elif kind == "gravity":
            a, b = np.asarray(g["field"], dtype=np.float64), _sub(o["field"], stride)
            err = np.abs(a - b).max()
            assert err <= max(field_tol, 3e-16), "%s err %g" % (tag, err)
        elif kind == "adding":
            np.testing.assert_allclose(np.asarray(o["particles_out"], dtype=np.float64),
                                       np.asarray(g["particles_out"], dtype=np.float64), rtol=0, atol=force_tol,
                                       err_msg=tag)
            np.testing.assert_array_equal(np.asarray(o["stable_out"]), np.asarray(g["stable_out"]), err_msg=tag)
        elif kind in ("force", "force_exp"):
            np.testing.assert_allclose(np.asarray(o["force"]), np.asarray(g["force"], dtype=np.float64), rtol=0,
                                       atol=force_tol, err_msg=tag)
        elif kind == "acc":
            for t in ("1", "2"):
                for key in ("TP", "FP", "FN"):
                    assert int(o["metrics"][t][key]) == int(g["metrics"][t][key]), "%s %s_%s" % (tag, t, key)
                for key in ("dice", "precision", "recall"):
                    assert float(o["metrics"][t][key]) == float(g["metrics"][t][key]), "%s %s_%s" % (tag, t, key)
        elif kind == "chamfer":
            for key in ("lsa_sum", "lsa_mean", "lsa_var"):
                np.testing.assert_allclose(float(o[key]), float(g[key]), rtol=1e-12, err_msg="%s %s" % (tag, key))
        elif kind == "blob_energy":
            np.testing.assert_allclose(np.asarray(o["energy"]), np.asarray(g["energy"], dtype=np.float64),
                                       rtol=0, atol=max(field_tol, 1e-15), err_msg=tag)
