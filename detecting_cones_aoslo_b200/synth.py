"""Synthetic AOSLO-like frames (SURVEY.md section 8d recipe): a jittered hexagonal cone lattice
rendered with a Gaussian PSF on a dark background, plus sensor noise, quantised to uint8.
Host-side input generation only (NumPy); nothing here is on the detection path."""
import numpy as np


def synth_aoslo(H, W, pitch=5.8, seed=2022):
    """-> (img uint8 (H,W), gt float64 (n,2) [x,y] in the CSV's 1-based convention)."""
    rng = np.random.default_rng(seed)
    p = float(pitch)
    row_h = p * np.sqrt(3.0) / 2.0
    ny, nx = int(H / row_h) + 2, int(W / p) + 2
    yy, xx = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64), indexing='ij')
    cx = (xx + 0.5 * (yy % 2)) * p
    cy = yy * row_h
    cx = cx + rng.normal(0.0, 0.08 * p, cx.shape)
    cy = cy + rng.normal(0.0, 0.08 * p, cy.shape)
    amp = rng.uniform(0.45, 1.0, cx.shape)
    keep = (cx >= 0) & (cx <= W - 1) & (cy >= 0) & (cy <= H - 1)
    cx, cy, amp = cx[keep], cy[keep], amp[keep]
    sig = 0.22 * p
    r = int(np.ceil(3.0 * sig))
    img = np.full((H, W), 0.18, dtype=np.float64)
    ix, iy = np.rint(cx).astype(np.int64), np.rint(cy).astype(np.int64)
    for dy in range(-r, r + 1):
        for dx in range(-r, r + 1):
            px, py = ix + dx, iy + dy
            ok = (px >= 0) & (px < W) & (py >= 0) & (py < H)
            val = amp[ok] * np.exp(-((px[ok] - cx[ok]) ** 2 + (py[ok] - cy[ok]) ** 2) / (2.0 * sig * sig))
            np.add.at(img, (py[ok], px[ok]), val * 0.62)
    img += rng.normal(0.0, 0.02, img.shape)
    img = np.clip(img, 0.0, 1.0)
    img8 = np.rint(img * 255.0).astype(np.uint8)
    gt = np.stack([cx + 1.0, cy + 1.0], axis=1)
    return img8, gt


def synth_config(H, W, pitch=5.8, boundary=15, **over):
    """The reference's `config_zoo` values (main_with_delitions.py:35-75) on a full H x W frame,
    with the pitch-dependent keys rescaled as SURVEY.md 8d prescribes (5.84 / 0.83 / 21 / 3.5 / 5
    at the shipped image's pitch)."""
    from .configs import zoo_config
    c = zoo_config()
    s = pitch / 5.84
    c.update({
        'region': {'x_min': 0, 'x_max': W - 1, 'y_min': 0, 'y_max': H - 1},
        'boundary_size': {'x': boundary, 'y': boundary},
        'mu_dist_func': 5.84 * s, 'sigma_dist_func': 0.83 * s,
        'reception_field_size': int(2 * np.ceil(1.8 * pitch) + 1) if abs(s - 1) > 0.05 else 21,
        'threshhold_dist_del': 3.5 * s, 'particle_call_window': max(2, int(round(5 * s))),
    })
    c.update(over)
    return c
