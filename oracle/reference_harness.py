"""TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.

Runs the UNMODIFIED reference modules from /root/reference (read-only) inside
this container so that their outputs can be recorded as golden vectors.  The
reference cannot be imported as is because three viz / image dependencies are
absent from the image; they are stubbed:

  matplotlib, matplotlib.pyplot, imageio  -> inert objects (visualisation only)
  skimage.feature                          -> oracle/skimage_shim.py (restated)
  wandb                                    -> inert object (never called: USE_WANDB=False)

Nothing here is available on the GPU box (/root/reference does not exist
there); tests read only the committed fixtures in tests/golden/ that
oracle/gen_golden.py produced with this harness.
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("AOSLO_REFERENCE_ROOT", "/root/reference")

_DATASET_STEM = "BAK1008L1_2020_07_02_11_56_18_AOSLO_788_V006_annotated_JLR_128_97_646_612"


class _Inert(types.ModuleType):
    """A module whose every attribute is a callable returning another inert object."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _InertCallable()


class _InertCallable:
    def __call__(self, *a, **k):
        return _InertCallable()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _InertCallable()

    def __iter__(self):
        return iter((_InertCallable(), _InertCallable()))

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def available():
    return os.path.isdir(REFERENCE_ROOT) and os.path.isfile(os.path.join(REFERENCE_ROOT, "utils.py"))


def load_reference():
    """Import the reference's utils / image_transformation / pipeline_with_delitions.

    Returns a namespace with attributes utils, image_transformation, pipeline.
    The modules are loaded under private names so that they never shadow the
    product package's modules of the same name.
    """
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from oracle import skimage_shim

    saved = {k: sys.modules.get(k) for k in
             ("matplotlib", "matplotlib.pyplot", "imageio", "skimage", "skimage.feature", "wandb",
              "utils", "image_transformation", "pipeline_with_delitions")}
    mpl = _Inert("matplotlib")
    plt = _Inert("matplotlib.pyplot")
    mpl.pyplot = plt
    sk = types.ModuleType("skimage")
    skf = types.ModuleType("skimage.feature")
    skf.hessian_matrix = skimage_shim.hessian_matrix
    skf.hessian_matrix_eigvals = skimage_shim.hessian_matrix_eigvals
    sk.feature = skf
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "imageio": _Inert("imageio"),
                        "skimage": sk, "skimage.feature": skf, "wandb": _Inert("wandb")})
    for k in ("utils", "image_transformation", "pipeline_with_delitions"):
        sys.modules.pop(k, None)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        it = importlib.import_module("image_transformation")
        ut = importlib.import_module("utils")
        pl = importlib.import_module("pipeline_with_delitions")
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in ("utils", "image_transformation", "pipeline_with_delitions"):
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            elif k in sys.modules and k not in ("utils", "image_transformation", "pipeline_with_delitions"):
                del sys.modules[k]
    ns = types.SimpleNamespace(utils=ut, image_transformation=it, pipeline=pl)
    return ns


def load_dataset():
    """The shipped image + GT exactly as the reference's __main__ reads them
    (pipeline_with_delitions.py:443-446: pd.read_csv eats CSV line 1 as a header)."""
    import cv2
    import pandas as pd
    d = os.path.join(REFERENCE_ROOT, "dataset")
    gt = pd.read_csv(os.path.join(d, _DATASET_STEM + ".csv")).to_numpy()
    img = cv2.imread(os.path.join(d, _DATASET_STEM + ".tiff"), -1)
    return img, gt


def closed_form_energy_pit_func2(x, y, mu, sigma, lambda_):
    """Closed form of utils.py:82-87 (scipy.stats.expon(loc=lambda_).pdf and
    scipy.stats.norm(mu, sigma).pdf written out).  Used only to make big traces
    affordable; SURVEY.md measured <= 1.1e-16 difference and identical runs."""
    dist = np.sqrt((x ** 2.) + (y ** 2.))
    z = dist - lambda_
    e = np.exp(-z) if z >= 0 else 0.0
    yv = (dist - mu) / sigma
    n = np.exp(-yv ** 2 / 2.0) / np.sqrt(2 * np.pi) / sigma
    if dist <= mu:
        return e - n + 0.48
    return e + n - 0.48


class Trace:
    """Records the arguments / results of the reference's stage functions during
    one find_blobs run by wrapping the names bound in the pipeline module."""

    def __init__(self):
        self.events = []

    def add(self, kind, **payload):
        self.events.append((kind, payload))

    def of(self, kind):
        return [p for k, p in self.events if k == kind]


def run_find_blobs(cfg, img=None, gt=None, trace=None, closed_form=False, quiet=True):
    """Run the reference's find_blobs (pipeline_with_delitions.py:12) and
    optionally trace its stage calls.  Returns (result, time_spended, trace)."""
    import contextlib
    import io

    ref = load_reference()
    pl, ut = ref.pipeline, ref.utils
    if img is None:
        img, gt = load_dataset()
    gt = np.array(gt, dtype=np.float64, copy=True)
    img_colorful = np.repeat(img[:, :, None], 3, axis=2)

    if closed_form:
        ut.energy_pit_func2 = closed_form_energy_pit_func2

    if trace is not None:
        def wrap(name, recorder):
            orig = getattr(pl, name)

            def w(*a, **k):
                out = orig(*a, **k)
                recorder(a, k, out)
                return out
            setattr(pl, name, w)

        wrap("calc_grad_field", lambda a, k, out: trace.add(
            "grad", R_b=np.array(a[0]), grad=np.array(out)))
        wrap("initialize_high_prop_location", lambda a, k, out: trace.add(
            "seeds", particles=np.array(out)))
        wrap("get_precomputed_dist_func", lambda a, k, out: trace.add(
            "lut", lut=np.array(out)))

        def rec_del(a, k, out):
            trace.add("delete", particles_in=np.array(a[0], dtype=np.float64),
                      stable_in=np.array(a[1], dtype=np.int8),
                      threshold=float(k.get("threshold_dist", 2.)),
                      particles_out=np.array(out[0]),
                      deleted=np.array([out[1][i] for i in range(len(out[1]))], dtype=bool),
                      stable_out=np.array(out[2], dtype=np.int8))
        # del_close_particles receives particles AFTER the in-place move step
        pl.del_close_particles_orig = pl.del_close_particles
        wrap("del_close_particles", rec_del)
        wrap("calc_gravity_field_optim", lambda a, k, out: trace.add(
            "gravity", particles=np.array(a[2], dtype=np.float64), field=np.array(out)))
        wrap("adding_particles", lambda a, k, out: trace.add(
            "adding", particles_out=np.array(out[0], dtype=np.float64),
            stable_out=np.array(out[1], dtype=np.int8)))

        # calc_dist_energy_pit_optim(particles, stable, n_neighbours=..., ...)
        def rec_force(a, k, out):
            trace.add("force", particles=np.array(a[0], dtype=np.float64),
                      stable=np.array(a[1], dtype=np.int8),
                      force=np.array(out[0]), modules=np.array(out[1]))
        wrap("calc_dist_energy_pit_optim", rec_force)
        wrap("calc_dist_energy", lambda a, k, out: trace.add(
            "force_exp", particles=np.array(a[0], dtype=np.float64),
            force=np.array(out[0]), modules=np.array(out[1])))

        def rec_acc(a, k, out):
            trace.add("acc", particles=np.array(a[0], dtype=np.float64), gt=np.array(a[1]),
                      gt_to_particles=np.array(a[2]), particles_to_gt=np.array(a[3]),
                      metrics={t: dict(v) for t, v in out.items()})
        wrap("calc_acc_metrics", rec_acc)
        wrap("calc_metrics", lambda a, k, out: trace.add(
            "chamfer", lsa_sum=float(out[1]), lsa_mean=float(out[2]), lsa_var=float(out[3])))
        wrap("calc_blob_energy", lambda a, k, out: trace.add(
            "blob_energy", energy=np.array(out)))

    sink = io.StringIO()
    ctx = contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()
    with ctx:
        result, time_spended = pl.find_blobs(cfg, gt, img, img_colorful, False, False)
    return result, time_spended, trace
