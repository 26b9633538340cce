"""Where the end-to-end rate goes: 8 lanes, K runs per lane back to back (bench.py's timed region), variants of the
host side of find_blobs."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from concurrent.futures import ThreadPoolExecutor
from detecting_cones_aoslo_b200.pipeline_with_delitions import find_blobs

L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = 10
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lanes = [bench.Lane(j, dev) for j in range(L)]
pool = ThreadPoolExecutor(max_workers=L)
main_s = torch.cuda.current_stream()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for ln in lanes:
    ln.d_raw = torch.from_numpy(ln.img).to(dev)


def region(fn):
    gts = {id(ln): ln.prep_e2e_steps(K) for ln in lanes}
    flush.fill_(1)
    torch.cuda._sleep(300000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main_s)
    for ln in lanes:
        ln.stream.wait_event(e0)
    t0 = time.perf_counter()

    def loop(ln):
        for i in range(K):
            fn(ln, gts[id(ln)][i])
    for f in [pool.submit(loop, ln) for ln in lanes]:
        f.result()
    for ln in lanes:
        d = torch.cuda.Event()
        d.record(ln.stream)
        main_s.wait_event(d)
    e1.record(main_s)
    torch.cuda.synchronize()
    return L * K / (e0.elapsed_time(e1) / 1e3)


def e2e(ln, gt):
    with torch.cuda.stream(ln.stream):
        find_blobs(dict(ln.cfg), gt, ln.img, None, False, False, engine=ln.eng)


def e2e_dev_img(ln, gt):
    with torch.cuda.stream(ln.stream):
        find_blobs(dict(ln.cfg), gt, ln.d_raw, None, False, False, engine=ln.eng)


def resident(ln, gt):
    ln.run_resident()


def resident_plus_python(ln, gt):
    # the resident run + the host-side work of find_blobs that does not touch the device
    from detecting_cones_aoslo_b200.pipeline_with_delitions import make_config_struct
    make_config_struct(ln.cfg, ln.H, ln.W)
    ln.run_resident()


for name, fn in (("resident", resident), ("resident + config packing", resident_plus_python), ("e2e", e2e),
                 ("e2e, frame already on the device (GT from the host)", e2e_dev_img), ("e2e", e2e),
                 ("resident", resident)):
    for _ in range(2):
        region(fn)
    print("%-55s %8.1f runs/s" % (name, np.mean([region(fn) for _ in range(3)])), flush=True)
