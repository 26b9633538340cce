"""Device engine: owns one libaoslo_b200 context and the torch tensors (device memory,
streams) the stage kernels work on.  PyTorch is plumbing here -- every computation is a
hand-written sm_100a kernel reached through the C ABI (include/aoslo_b200.h)."""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from ._lib import lib, check


def gaussian_half_kernel(sigma, truncate=4.0):
    """scipy.ndimage._filters._gaussian_kernel1d (order 0) -> (half weights w[0..r], r).
    Parameter set-up only (9 floats for sigma = 1); the convolution runs on the device."""
    sd = float(sigma)
    radius = int(truncate * sd + 0.5)
    sigma2 = sd * sd
    x = np.arange(-radius, radius + 1)
    phi_x = np.exp(-0.5 / sigma2 * x ** 2)
    phi_x = phi_x / phi_x.sum()
    return phi_x[radius:].astype(np.float64), radius


def _require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("detecting_cones_aoslo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


class Engine:
    """One context per (device, padded image shape, capacities)."""

    def __init__(self, H, W, particle_capacity=None, gt_capacity=65536, device=None):
        _require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.H, self.W = int(H), int(W)
        if particle_capacity is None:
            particle_capacity = self.H * self.W          # seeds can cover every inner pixel
        self.cap = int(max(2, particle_capacity))
        self.gt_cap = int(max(1, gt_capacity))
        h = C.c_void_p()
        check(lib.aoslo_ctx_create(C.byref(h), self.device.index, self.H, self.W, self.cap, self.gt_cap),
              "aoslo_ctx_create")
        self._h = h
        self._n_out = C.c_int(0)
        self._fields = {}
        self._in_bufs = None
        self._raw_bufs = None

    def set_blocking_sync(self, enable=True):
        """Host waits sleep instead of spinning: for many concurrent lanes (one host thread each)."""
        check(lib.aoslo_ctx_set_blocking_sync(self._h, 1 if enable else 0), "aoslo_ctx_set_blocking_sync")

    def set_tuning(self, **knobs):
        """Implementation knobs of the library (k1_tma, pair_th, pair_occ, dtile_th, k1_kernel, no_graph);
        results never depend on them."""
        for k, v in knobs.items():
            check(lib.aoslo_ctx_set_tuning(self._h, k.encode(), int(v)), "aoslo_ctx_set_tuning(%s)" % k)

    def close(self):
        if getattr(self, "_h", None):
            lib.aoslo_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def dev(self, a, dtype):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(self.device)

    def upload_image(self, img):
        """uint8 (H,W) host image -> device tensor VIEW (H,W) whose row pitch is a multiple of 16 bytes,
        so that the blobness kernel can fetch 4 pixels per 32-bit load."""
        img = np.ascontiguousarray(img)
        H, W = img.shape
        pitch = (W + 15) // 16 * 16
        buf = torch.zeros((H, pitch), dtype=torch.uint8, device=self.device)
        view = buf[:, :W]
        view.copy_(torch.as_tensor(img), non_blocking=False)
        return view

    def prepare_inputs(self, img, gt, region, boundary, sync=True, gt_is_raw=False):
        """Device-side np.round + crop_image + mirror_padding_image (reference
        pipeline_with_delitions.py:27-33).  img: (H0,W0) uint8 host array (or a device uint8 tensor);
        gt: (n,2) float64 host array -- gt_is_raw = False: ALREADY shifted by -1 by the caller (the reference's
        in-place side effect); gt_is_raw = True: the caller's own 1-based array, which must be float64,
        C-contiguous and writable: the device copy is shifted by the prepare kernel and the in-place shift of the
        host array is deferred into the next run() (aoslo_stage_gt), or flush_host_work().  Returns the padded
        image view (pitch multiple of 16) and the cropped / shifted GT, both on the device."""
        on_device = isinstance(img, torch.Tensor)
        if not on_device:
            img = np.asarray(img)
        if img.dtype not in (np.uint8, torch.uint8) or img.ndim != 2:
            # the kernels take the raw bytes; the reference's skimage path would rescale other dtypes
            # (img_as_float), which is not implemented -- refuse instead of reinterpreting the buffer
            raise TypeError("img must be a 2-D uint8 array (got %s, ndim %d)" % (img.dtype, img.ndim))
        if on_device:
            if not img.is_cuda or img.device != self.device:
                raise ValueError("a tensor image must live on the engine's device (dataio.PinnedBatch)")
            img = img.contiguous()
        else:
            img = np.ascontiguousarray(img)
        H0, W0 = img.shape
        assert (0 <= region['x_min']) and (0 <= region['y_min'])                       # image_transformation.py:7
        assert (W0 > region['x_max']) and (H0 > region['y_max'])                       # image_transformation.py:8
        if gt_is_raw:
            g = gt
            if not (isinstance(g, np.ndarray) and g.dtype == np.float64 and g.flags.c_contiguous
                    and g.flags.writeable and g.ndim == 2 and g.shape[1] == 2):
                raise TypeError("gt_is_raw needs a writable C-contiguous (n,2) float64 array")
        else:
            g = np.ascontiguousarray(gt, dtype=np.float64).reshape(-1, 2)
        if g.shape[0] > self.gt_cap:
            raise MemoryError("GT capacity of the context exceeded")
        # context-owned device staging (stable addresses, no allocation per call); pageable host arrays go
        # through cudaMemcpyAsync's own staging, pinned ones (torch pin_memory) are true async copies
        if self._raw_bufs is None or self._raw_bufs[0].numel() < H0 * W0:
            self._raw_bufs = (self.empty((H0 * W0,), torch.uint8), self.empty((self.gt_cap, 2), torch.float64))
        stream = self._stream()
        if on_device:
            d_raw = img                                    # already uploaded (batched pinned upload)
        else:
            d_raw = self._raw_bufs[0][: H0 * W0].view(H0, W0)
            d_raw.copy_(torch.from_numpy(img), non_blocking=True)
        d_gt_raw = self._raw_bufs[1]
        check(lib.aoslo_stage_gt(self._h, stream, g.ctypes.data, g.shape[0], d_gt_raw.data_ptr(),
                                 1 if gt_is_raw else 0), "aoslo_stage_gt")
        pitch = (self.W + 15) // 16 * 16
        if self._in_bufs is None:          # context-owned: stable addresses (CUDA-graph replay in aoslo_run)
            self._in_bufs = (self.empty((self.H, pitch), torch.uint8), self.empty((self.gt_cap, 2), torch.float64))
        buf, d_gt = self._in_bufs
        check(lib.aoslo_prepare_inputs(self._h, stream, self._p(d_raw), H0, W0, W0,
                                       int(region['x_min']), int(region['x_max']), int(region['y_min']),
                                       int(region['y_max']), int(boundary['x']), int(boundary['y']),
                                       self._p(buf), pitch, d_gt_raw.data_ptr(), g.shape[0], d_gt.data_ptr(),
                                       C.byref(self._n_out) if sync else None, 1 if gt_is_raw else 0),
              "aoslo_prepare_inputs")
        if not sync:
            # the cropped GT count stays on the device (aoslo_run reads it there): no host synchronisation;
            # the returned view has the RAW length, only its first <count> rows are data
            return buf[:, :self.W], d_gt[: g.shape[0]]
        return buf[:, :self.W], d_gt[: self._n_out.value]

    def flush_host_work(self):
        """performs a deferred in-place GT shift now (after the upload) if no run() has picked it up"""
        check(lib.aoslo_ctx_flush_host_work(self._h), "aoslo_ctx_flush_host_work")

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    # FIELD buffers (R_b (H,W), gradient (H,W,2)): the library stores their rows with a pitch of fpitch(W) pixels
    # (W rounded up to a multiple of 4: 16-byte aligned rows for the TMA stores of the blobness kernel).  Tensors
    # are VIEWS [:, :W] of the padded buffers; .cpu() / .contiguous() give the dense array.
    @staticmethod
    def fpitch(W):
        return (int(W) + 3) // 4 * 4

    def field(self, H, W, dtype, components=1):
        if components == 1:
            return self.empty((H, self.fpitch(W)), dtype)[:, :W]
        return self.empty((H, self.fpitch(W), components), dtype)[:, :W, :]

    def to_field(self, a, dtype):
        """dense (H,W) / (H,W,2) host array or device tensor -> padded device field view."""
        t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
        out = self.field(t.shape[0], t.shape[1], dtype, 1 if t.dim() == 2 else t.shape[2])
        out.copy_(t.to(device=self.device, dtype=dtype))
        return out

    def _pf(self, t):
        """pointer of a field view; a dense tensor whose width is not a multiple of 4 is a caller bug."""
        if t is None:
            return C.c_void_p(0)
        per_px = 1 if t.dim() == 2 else t.shape[2]
        if t.stride(0) != self.fpitch(t.shape[1]) * per_px or t.stride(1) != per_px:
            raise ValueError("field tensors must come from Engine.field / Engine.to_field (padded row pitch)")
        return C.c_void_p(t.data_ptr())

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def status(self, clear=True):
        w = C.c_int(0)
        check(lib.aoslo_ctx_status(self._h, self._stream(), C.byref(w), 1 if clear else 0), "aoslo_ctx_status")
        return w.value

    def clear_status(self):
        """Zero the status word, stream ordered (no host synchronisation)."""
        check(lib.aoslo_ctx_clear_status(self._h, self._stream()), "aoslo_ctx_clear_status")

    def raise_on_status(self):
        w = self.status(clear=True)
        if w:
            _lib.raise_device_status(w)

    @staticmethod
    def _fdt(t):
        if t.dtype == torch.float32:
            return _lib.F32
        if t.dtype == torch.float64:
            return _lib.F64
        raise TypeError("field must be float32 or float64, got %s" % t.dtype)

    # ------------------------------------------------------------------ L2 field
    _SCALE_CACHE = {}

    @classmethod
    def _scale_tables(cls, scales):
        """Per-scale Gaussian half kernels / radii / sigma^2 (host parameter tables, cached)."""
        hit = cls._SCALE_CACHE.get(scales)
        if hit is not None:
            return hit
        ns = len(scales)
        if ns < 1 or ns > _lib.MAX_SCALES:
            raise ValueError("1..%d scales" % _lib.MAX_SCALES)
        wts = np.zeros((ns, _lib.MAX_RADIUS + 1), dtype=np.float64)
        radii = np.zeros(ns, dtype=np.int32)
        s2 = np.zeros(ns, dtype=np.float64)
        for i, s in enumerate(scales):
            w, r = gaussian_half_kernel(s)
            if r > _lib.MAX_RADIUS:
                raise NotImplementedError("sigma %.2f needs radius %d > %d" % (s, r, _lib.MAX_RADIUS))
            wts[i, :r + 1] = w
            radii[i] = r
            s2[i] = float(s) * float(s)
        cls._SCALE_CACHE[scales] = (wts, radii, s2)
        return wts, radii, s2

    def blobness(self, img_u8, formula='custom', grad_type='np_grad', scales=(1.0,), dtype=torch.float64,
                 want_grad=True, out=None):
        """u8 (H,W) device tensor -> (R_b (H,W), grad (H,W,2)) of `dtype`."""
        if formula == 'div_corrected':
            raise IndexError("blobness_formula='div_corrected' is broken in the reference "
                             "(pipeline_with_delitions.py:77, list-of-mask indexing)")
        if formula not in _lib.FORMULA:
            raise NotImplementedError
        if grad_type not in _lib.GRAD:
            raise AttributeError('grad_type {0} is not implemented'.format(grad_type))
        assert img_u8.dtype == torch.uint8 and img_u8.is_cuda and img_u8.dim() == 2
        H, W = img_u8.shape
        ns = len(scales)
        wts, radii, s2 = self._scale_tables(tuple(float(s) for s in scales))
        if out is None:
            Rb = self.field(H, W, dtype)
            grad = self.field(H, W, dtype, 2) if want_grad else None
        else:
            Rb, grad = out
        check(lib.aoslo_blobness(self._h, self._stream(), self._p(img_u8), H, W, img_u8.stride(0),
                                 wts.ctypes.data_as(C.POINTER(C.c_double)),
                                 radii.ctypes.data_as(C.POINTER(C.c_int)),
                                 s2.ctypes.data_as(C.POINTER(C.c_double)), ns,
                                 _lib.FORMULA[formula], _lib.GRAD[grad_type], self._fdt(Rb),
                                 self._pf(Rb), self._pf(grad)), "aoslo_blobness")
        return Rb, grad

    def grad_field(self, field, grad_type):
        if grad_type not in _lib.GRAD:
            raise AttributeError('grad_type {0} is not implemented'.format(grad_type))
        H, W = field.shape
        out = self.field(H, W, field.dtype, 2)
        check(lib.aoslo_grad_field(self._h, self._stream(), self._pf(field), H, W, _lib.GRAD[grad_type],
                                   self._fdt(field), self._pf(out)), "aoslo_grad_field")
        return out

    # ------------------------------------------------------------------ L3 particle system
    def seed(self, Rb, bx, by, threshold):
        pos = self.empty((self.cap, 2), torch.float64)
        check(lib.aoslo_seed(self._h, self._stream(), self._pf(Rb), self._fdt(Rb), self.H, self.W, int(bx), int(by),
                             float(threshold), self._p(pos), self.cap, C.byref(self._n_out)), "aoslo_seed")
        return pos[: self._n_out.value]

    def knn(self, points, query, k, tie_mode='canonical'):
        n, nq = points.shape[0], query.shape[0]
        if k > n:
            raise ValueError("Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %d" % (k, n))
        idx = self.empty((nq, k), torch.int32)
        dist = self.empty((nq, k), torch.float64)
        check(lib.aoslo_knn(self._h, self._stream(), self._p(points), n, self._p(query), nq, int(k),
                            _lib.TIE[tie_mode], self._p(idx), self._p(dist)), "aoslo_knn")
        return dist, idx

    def delete_close(self, pos, stable, Rb, threshold, tie_mode='canonical'):
        n = pos.shape[0]
        if n < 2:
            raise ValueError("Expected n_neighbors <= n_samples_fit, but n_neighbors = 2, n_samples_fit = %d" % n)
        pos_out = self.empty((n, 2), torch.float64)
        st_out = self.empty((n,), torch.uint8)
        deleted = self.empty((n,), torch.uint8)
        check(lib.aoslo_delete_close(self._h, self._stream(), self._p(pos), self._p(stable), n, self._pf(Rb),
                                     self._fdt(Rb), self.H, self.W, float(threshold), _lib.TIE[tie_mode],
                                     self._p(pos_out), self._p(st_out), self._p(deleted), C.byref(self._n_out)),
              "aoslo_delete_close")
        m = self._n_out.value
        return pos_out[:m], st_out[:m], deleted

    def dist_lut(self, rfs, mu, sigma, lambda_):
        lut = self.empty((rfs, rfs), torch.float64)
        check(lib.aoslo_dist_lut(self._h, self._stream(), int(rfs), float(mu), float(sigma), float(lambda_),
                                 self._p(lut)), "aoslo_dist_lut")
        return lut

    def gravity_field(self, pos, lut):
        field = self.empty((self.H, self.W), torch.float64)
        check(lib.aoslo_gravity_field(self._h, self._stream(), self._p(pos), pos.shape[0], self._p(lut),
                                      lut.shape[0], self.H, self.W, self._p(field)), "aoslo_gravity_field")
        return field

    def add_particles(self, field, pos, stable, bx, by, window, threshold):
        if not isinstance(window, (int, np.integer)):
            # reference: range(..., step) with a float step -> TypeError (utils.py:341)
            raise TypeError("'%s' object cannot be interpreted as an integer" % type(window).__name__)
        n = pos.shape[0]
        nwin = (math.ceil((self.W - 2 * bx) / window)) * (math.ceil((self.H - 2 * by) / window))
        cap = min(self.cap, n + nwin)
        pos2 = self.empty((cap, 2), torch.float64)
        st2 = self.empty((cap,), torch.uint8)
        pos2[:n] = pos
        st2[:n] = stable
        check(lib.aoslo_add_particles(self._h, self._stream(), self._p(field), self.H, self.W, int(bx), int(by),
                                      int(window), float(threshold), self._p(pos2), self._p(st2), n, cap,
                                      C.byref(self._n_out)), "aoslo_add_particles")
        m = self._n_out.value
        return pos2[:m], st2[:m]

    def dist_force(self, pos, stable, n_neighbours, energy_type, p0, p1, p2=0.0, tie_mode='canonical'):
        n = pos.shape[0]
        if n_neighbours + 1 > n:
            raise ValueError("Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %d"
                             % (n_neighbours + 1, n))
        force = self.empty((n, 2), torch.float64)
        check(lib.aoslo_dist_force(self._h, self._stream(), self._p(pos), self._p(stable), n, int(n_neighbours),
                                   _lib.ENERGY[energy_type], float(p0), float(p1), float(p2), _lib.TIE[tie_mode],
                                   self._p(force)), "aoslo_dist_force")
        return force

    def move(self, pos, stable, grad, force, lambda_blob, lambda_dist, step_mode):
        check(lib.aoslo_move(self._h, self._stream(), self._p(pos), self._p(stable), pos.shape[0], self._pf(grad),
                             self._fdt(grad), self.H, self.W, self._p(force), float(lambda_blob), float(lambda_dist),
                             _lib.STEP[step_mode]), "aoslo_move")
        return pos

    # ------------------------------------------------------------------ L4 metrics
    def metrics(self, pos, stable, gt, Rb, bx, by, want_dists=False):
        n, g = pos.shape[0], gt.shape[0]
        rec = _lib.MetricsRecord()
        p2g = self.empty((n,), torch.float64) if want_dists else None
        g2p = self.empty((g,), torch.float64) if want_dists else None
        check(lib.aoslo_metrics(self._h, self._stream(), self._p(pos), self._p(stable), n, self._p(gt), g,
                                self._pf(Rb), self._fdt(Rb), self.H, self.W, int(bx), int(by), C.byref(rec),
                                self._p(p2g), self._p(g2p)), "aoslo_metrics")
        return rec, p2g, g2p

    def hungarian(self, gt, pos, want_costs=False):
        """calc_metrics(mode='hangarian') (reference utils.py:228-234): distance matrix + scipy's rectangular
        assignment solver restated on the device.  Returns (sum, mean, var, assigned costs or None, steps)."""
        g, n = gt.shape[0], pos.shape[0]
        out = (C.c_double * 4)()
        sel = self.empty((min(g, n),), torch.float64) if want_costs else None
        check(lib.aoslo_hungarian(self._h, self._stream(), self._p(gt), g, self._p(pos), n, out, self._p(sel)),
              "aoslo_hungarian")
        return out[0], out[1], out[2], sel, int(out[3])

    def classify(self, g2p, p2g, pos, bx, by, threshold):
        """TP / FN flags per GT point and FP flags per particle (reference utils.visualize_all)."""
        gt_tp = self.empty((g2p.shape[0],), torch.uint8)
        p_fp = self.empty((pos.shape[0],), torch.uint8)
        check(lib.aoslo_classify(self._h, self._stream(), self._p(g2p), g2p.shape[0], self._p(p2g), self._p(pos),
                                 pos.shape[0], int(bx), int(by), self.H, self.W, float(threshold), self._p(gt_tp),
                                 self._p(p_fp)), "aoslo_classify")
        return gt_tp, p_fp

    def iterate(self, cfg_struct, Rb, grad, pos, stable):
        """one epoch body: force + move + deletion -> (pos, stable) (reference :173-253)."""
        n = pos.shape[0]
        pos_out = self.empty((n, 2), torch.float64)
        st_out = self.empty((n,), torch.uint8)
        check(lib.aoslo_iterate(self._h, self._stream(), C.byref(cfg_struct), self._pf(Rb), self._pf(grad),
                                self._fdt(Rb), self._p(pos), self._p(stable), n, self._p(pos_out), self._p(st_out),
                                C.byref(self._n_out)), "aoslo_iterate")
        m = self._n_out.value
        return pos_out[:m], st_out[:m]

    # ------------------------------------------------------------------ L5 whole run
    def sweep(self, cfg_structs, Rb, grad, gt):
        """aoslo_sweep: the configurations of `cfg_structs` (list of _lib.Config) back to back on one field ->
        list of _lib.TrialResult."""
        n = len(cfg_structs)
        cfgs = (_lib.Config * max(n, 1))(*cfg_structs)
        res = (_lib.TrialResult * max(n, 1))()
        check(lib.aoslo_sweep(self._h, self._stream(), cfgs, n, self._pf(Rb), self._pf(grad), self._fdt(Rb),
                              self._p(gt), gt.shape[0], res), "aoslo_sweep")
        return [res[i] for i in range(n)]

    def field_buffers(self, dtype):
        """Context-owned (H,W) / (H,W,2) field buffers: stable addresses let aoslo_run replay its cached
        CUDA graph instead of re-capturing it."""
        key = str(dtype)
        if key not in self._fields:
            self._fields[key] = (self.field(self.H, self.W, dtype), self.field(self.H, self.W, dtype, 2))
        return self._fields[key]

    def run(self, cfg_struct, Rb, grad, gt, max_records=None, collect_times=True, n_gt_on_device=False):
        """aoslo_run.  The status word is sticky: call clear_status() before the field computation of a new run
        (find_blobs does); summ.status then carries the field's eigenvalue assert as well as the run's conditions.
        collect_times=False skips the five stage timings (one-thread time-stamp kernels in the run graph)."""
        if max_records is None:
            max_records = cfg_struct.epoch_num // max(1, cfg_struct.metric_measure_freq) + 2
        recs = (_lib.MetricsRecord * max_records)()
        summ = _lib.RunSummary()
        ms = (C.c_float * 5)() if collect_times else None
        check(lib.aoslo_run(self._h, self._stream(), C.byref(cfg_struct), self._pf(Rb), self._pf(grad), self._fdt(Rb),
                            self._p(gt), -1 if n_gt_on_device else gt.shape[0], recs, max_records, None, None,
                            C.byref(summ), ms),
              "aoslo_run")
        # the final particles are copied out AFTER the run, sized by its result (no capacity-sized output)
        n = summ.n_particles
        pos = self.empty((n, 2), torch.float64)
        st = self.empty((n,), torch.uint8)
        if n:
            check(lib.aoslo_read_particles(self._h, self._stream(), self._p(pos), self._p(st), n),
                  "aoslo_read_particles")
        return summ, [recs[i] for i in range(summ.n_records)], pos, st, (list(ms) if ms is not None else [0.0] * 5)
