"""Times the blobness kernels alone (CUDA events, L2 flushed between launches, host kept ahead of the device)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from detecting_cones_aoslo_b200.pipeline_with_delitions import prepare_inputs
from detecting_cones_aoslo_b200.utils import get_engine

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
img, gt, cfg = bench.make_workload(0, size)
pimg, pgt = prepare_inputs(cfg, gt.copy(), img)
H, W = pimg.shape
eng = get_engine(H, W, n_gt=pgt.shape[0])
d_img = eng.upload_image(pimg)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=eng.device)
out = {}
for name, dt in (("f32", torch.float32), ("f64", torch.float64)):
    Rb = eng.field(H, W, dt); g = eng.field(H, W, dt, 2)
    ms = []
    for i in range(13):
        flush.fill_(1); torch.cuda._sleep(300000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.blobness(d_img, 'custom', 'np_grad', (1.0,), dt, out=(Rb, g)); b.record()
        torch.cuda.synchronize()
        if i >= 3: ms.append(a.elapsed_time(b))
    bpp = 13.0 if dt == torch.float32 else 25.0
    out[name] = {"ms": float(np.mean(ms)), "min_ms": float(np.min(ms)), "GB/s": bpp * H * W / np.mean(ms) / 1e6}
print(json.dumps({"H": H, "W": W, "chunk": os.environ.get("AOSLO_MARCH_CHUNK"), **out}))
