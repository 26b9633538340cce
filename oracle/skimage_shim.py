"""TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.

Restatement of the two scikit-image functions the reference calls
(pipeline_with_delitions.py:36 and :38).  scikit-image is an un-vendored,
un-pinned third-party dependency of the reference and is NOT installed in this
image, so its published algorithm (0.19.x era, matching the reference's 2022
vintage) is restated here:

  hessian_matrix(image, sigma, mode, cval, order='rc'):
      image  = img_as_float(image)            (uint8 -> float64 / 255)
      G      = scipy.ndimage.gaussian_filter(image, sigma, mode=mode, cval=cval)
      g      = np.gradient(G)
      H      = [d(g[0])/dr, d(g[0])/dc, d(g[1])/dc]      (np.gradient again)
  hessian_matrix_eigvals([Hrr, Hrc, Hcc]):
      closed form for 2x2 symmetric matrices, descending order:
      (Hrr+Hcc)/2 +- sqrt(Hrc^2 + ((Hrr-Hcc)/2)^2)

Parity for this boundary is UNPINNED by the reference (it ships no tests); the
pin is this restatement + the goldens generated through it.
"""
from itertools import combinations_with_replacement

import numpy as np
from scipy import ndimage as ndi


def img_as_float(image):
    image = np.asarray(image)
    if image.dtype == np.uint8:
        return image.astype(np.float64) / 255.0
    if image.dtype.kind == "f":
        return image.astype(np.float64, copy=False)
    raise TypeError("oracle shim supports uint8 / float images only, got %s" % image.dtype)


def hessian_matrix(image, sigma=1, mode="constant", cval=0, order="rc"):
    image = img_as_float(image)
    gaussian_filtered = ndi.gaussian_filter(image, sigma=sigma, mode=mode, cval=cval)
    gradients = np.gradient(gaussian_filtered)
    axes = range(image.ndim)
    if order == "xy":
        axes = reversed(axes)
    return [np.gradient(gradients[ax0], axis=ax1)
            for ax0, ax1 in combinations_with_replacement(axes, 2)]


def hessian_matrix_eigvals(H_elems):
    M00, M01, M11 = H_elems
    eigs = np.empty((2,) + M00.shape, M00.dtype)
    eigs[:] = (M00 + M11) / 2
    hsqrtdet = np.sqrt(M01 ** 2 + ((M00 - M11) / 2) ** 2)
    eigs[0] += hsqrtdet
    eigs[1] -= hsqrtdet
    return eigs
