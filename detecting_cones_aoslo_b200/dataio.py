"""Input path of the detection run: TIFF / CSV readers and the batched pinned upload.

Replaces `cv2.imread(path, -1)` + `pd.read_csv(path).to_numpy()` of the reference's drivers
(pipeline_with_delitions.py:443-446, main_with_delitions.py:153-155).  The decoders are native (csrc/io_host.cpp,
through the C ABI) and write straight into pinned host memory, so a batch of frames reaches the device with ONE
asynchronous copy; `find_blobs` accepts the resulting device images as they are.

Header-row quirk: the annotation CSV has no header line, and `pd.read_csv` consumes its first line as column names,
so the reference silently drops the first ground-truth point (7 627 of 7 628 on the shipped image).
`header_row_quirk=True` (default) reproduces that; False keeps every point.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import lib, check


def tiff_shape(path):
    H, W, bits, samples = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    check(lib.aoslo_tiff_info(os.fsencode(path), C.byref(H), C.byref(W), C.byref(bits), C.byref(samples)),
          "aoslo_tiff_info(%s)" % path)
    return H.value, W.value, bits.value, samples.value


def read_tiff(path, out=None):
    """8-bit grayscale TIFF -> uint8 (H, W) array (`out`: a C-contiguous uint8 buffer to decode into)."""
    H, W, bits, samples = tiff_shape(path)
    if out is None:
        out = np.empty((H, W), dtype=np.uint8)
    if out.shape != (H, W) or out.dtype != np.uint8 or not out.flags['C_CONTIGUOUS']:
        raise ValueError("out must be a C-contiguous uint8 array of shape (%d, %d)" % (H, W))
    check(lib.aoslo_tiff_read_gray8(os.fsencode(path), out.ctypes.data_as(C.c_void_p), W),
          "aoslo_tiff_read_gray8(%s)" % path)
    return out


def read_points_csv(path, header_row_quirk=True):
    """Annotation CSV ("x,y" per line, 1-based) -> float64 (n, 2)."""
    n = C.c_int(0)
    st = lib.aoslo_csv_read_points(os.fsencode(path), 1 if header_row_quirk else 0, None, 0, C.byref(n))
    if st not in (_lib.OK, _lib.ERR_CAPACITY):
        check(st, "aoslo_csv_read_points(%s)" % path)
    out = np.empty((n.value, 2), dtype=np.float64)
    if n.value:
        check(lib.aoslo_csv_read_points(os.fsencode(path), 1 if header_row_quirk else 0,
                                        out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)),
              "aoslo_csv_read_points(%s)" % path)
    return out


def load_pair(img_path, csv_path, header_row_quirk=True):
    """(img uint8 (H,W), GT float64 (n,2)) exactly as the reference's drivers hand them to find_blobs."""
    return read_tiff(img_path), read_points_csv(csv_path, header_row_quirk)


class PinnedBatch:
    """Decodes a list of (tiff, csv) pairs into ONE pinned staging buffer and uploads all frames with one
    asynchronous host->device copy.  frames[i] = (device uint8 (H,W) tensor, GT float64 (n,2) host array); the
    device images are views of one allocation and can be passed to find_blobs as `img`."""

    def __init__(self, pairs, device=None, header_row_quirk=True):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("detecting_cones_aoslo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        shapes = [tiff_shape(p)[:2] for p, _ in pairs]
        offs, total = [], 0
        for H, W in shapes:
            offs.append(total)
            total += (H * W + 255) // 256 * 256            # 256-byte aligned frames
        self.host = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        hview = self.host.numpy()
        self.gt = []
        for (p, c), (H, W), o in zip(pairs, shapes, offs):
            read_tiff(p, out=hview[o:o + H * W].reshape(H, W))
            self.gt.append(read_points_csv(c, header_row_quirk))
        self.dev = torch.empty_like(self.host, device=self.device)
        self.dev.copy_(self.host, non_blocking=True)       # one copy for the whole batch
        self.h2d_bytes = int(total)
        self.frames = [(self.dev[o:o + H * W].view(H, W), g) for (H, W), o, g in zip(shapes, offs, self.gt)]

    def __len__(self):
        return len(self.frames)

    def __getitem__(self, i):
        return self.frames[i]
