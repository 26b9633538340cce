"""cfg5 / cfg4 on N ranks: the table gathered with one NCCL all-gather must equal the table one rank computes alone,
bit for bit in every column but the wall-clock one (SURVEY 8e verification).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sweep_ranks.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from detecting_cones_aoslo_b200 import sweep
from detecting_cones_aoslo_b200.synth import synth_aoslo, synth_config

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
img, gt = synth_aoslo(768, 768, 5.8, seed=2022)
base = synth_config(768, 768, 5.8, boundary=15, tie_mode='canonical', collect_stage_times=False)
n = 48
runner = sweep.UnitRunner(device=local, lanes=4)


def trials(idx):
    out = []
    for i in idx:
        c = sweep.sample_trial(base, i)
        c['region'] = dict(base['region'])
        c['tie_mode'] = 'canonical'
        out.append((i, c))
    return out


mine = runner.run_trials(img, gt, trials(sweep.shard_units(n, rank, world)))
table = sweep.gather_records(mine, n, device=dev)
frames = [(synth_aoslo(384, 384, 5.8, seed=2022 + i) + (synth_config(384, 384, 5.8, boundary=15, epoch_num=8,
                                                                      tie_mode='canonical',
                                                                      collect_stage_times=False),)) for i in range(12)]
tb = sweep.run_batch_of_frames(frames, rank, world, dev, runner=runner)
cols = [i for i, f in enumerate(sweep.RECORD_FIELDS) if f != 'time_spended']
if rank == 0:
    alone = np.array(runner.run_trials(img, gt, trials(range(n))))
    alone_b = np.array(runner.run_frames([(i,) + tuple(fr) for i, fr in enumerate(frames)]))
    ok = np.array_equal(table[:, cols], alone[:, cols]) and np.array_equal(tb[:, cols], alone_b[:, cols])
    print("ranks %d: sweep table %s, batch table %s, equal to the 1-rank tables: %s; failed units %d"
          % (world, table.shape, tb.shape, ok, int((table[:, 1] != 0).sum() + (tb[:, 1] != 0).sum())))
    assert ok
runner.close()
dist.barrier()
dist.destroy_process_group()
