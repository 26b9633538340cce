"""Summarises ncu outputs brought back in gpurun_out/ into the tracked profiles/ directory.

    python scripts/summarise_profiles.py <launch_csv> <round_tag> [<ncu-rep> ...]
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch_summary(path, tag):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    names = [(r['Kernel Name'], float(r['Metric Value'].replace(',', ''))) for r in rows]
    idx = [i for i, (n, _) in enumerate(names) if 'blobness' in n]
    # step boundaries: a blobness launch starts every run; take the LAST complete resident run before e2e
    start, end = idx[1], idx[2]
    agg = collections.OrderedDict()
    for n, v in names[start:end]:
        key = re.sub(r'\(.*', '', n)
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    out = ["# %s: kernel launch list of one detection run (ncu --metrics gpu__time_duration.sum, cold cache, serialised)" % tag,
           "", "command: `python bench.py --steps 1 --warmup 1 --no-cpu --batch 1` (2048^2 synthetic frame, M = 30)",
           "launches in the run: %d, summed kernel time: %.1f us" % (end - start, tot / 1e3), "",
           "| kernel | launches | total us | share | avg us | max us |", "|---|---:|---:|---:|---:|---:|"]
    for k, (c, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f%% | %.2f | %.2f |" % (k.replace('void ', '').replace('aoslo::', '')[:70],
                                                                    c, t / 1e3, 100 * t / tot, t / 1e3 / c, m / 1e3))
    return "\n".join(out) + "\n"


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__waves_per_multiprocessor', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'sm__cycles_active.avg', 'sm__cycles_elapsed.max', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active']


def rep_summary(path):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out, facts = [], {}
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')]
        out.append("## `%s` (%s)" % (name, os.path.basename(path)))
        out.append("")
        out.append("| metric | value | unit |")
        out.append("|---|---:|---|")
        for w in WANT:
            if w in hdr:
                out.append("| %s | %s | %s |" % (w, r[hdr.index(w)], units[hdr.index(w)]))
                facts[w] = (r[hdr.index(w)], units[hdr.index(w)])
        st = [(float(r[i]), h) for i, h in enumerate(hdr)
              if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')
              and r[i] not in ('', 'n/a')]
        out.append("")
        out.append("top stall reasons (warps stalled per issued instruction): " + ", ".join(
            "%s %.2f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v)
            for v, h in sorted(st, reverse=True)[:6]))
        out.append("")
    return "\n".join(out), facts


def to_bytes(val, unit):
    v = float(val.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


if __name__ == '__main__':
    launch_csv, tag = sys.argv[1], sys.argv[2]
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    with open(os.path.join(ROOT, 'profiles', '%s_launches.md' % tag), 'w') as f:
        f.write(launch_summary(launch_csv, tag))
    parts, traffic = ["# %s: ncu --set full --clock-control none captures of the fused blobness kernel (K1)\n" % tag], {}
    others = ["# %s: ncu --set full --clock-control none captures of particle-system kernels (one launch each, "
              "2048^2 synthetic frame)\n" % tag]
    for rep in sys.argv[3:]:
        text, facts = rep_summary(rep)
        if 'k1' not in os.path.basename(rep):
            others.append(text)
            continue
        parts.append(text)
        if 'dram__bytes_read.sum' in facts:
            key = 'f64' if 'f64' in rep else 'f32'
            traffic[key] = to_bytes(*facts['dram__bytes_read.sum']) + to_bytes(*facts['dram__bytes_write.sum'])
    with open(os.path.join(ROOT, 'profiles', '%s_k1_ncu.md' % tag), 'w') as f:
        f.write("\n".join(parts))
    if len(others) > 1:
        with open(os.path.join(ROOT, 'profiles', '%s_particles_ncu.md' % tag), 'w') as f:
            f.write("\n".join(others))
    with open(os.path.join(ROOT, 'profiles', '%s_k1_traffic.json' % tag), 'w') as f:
        json.dump({"dram_bytes_per_launch": traffic,
                   "note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch on the 2077x2077 frame; "
                           "writes that still sit in the 126 MB L2 at kernel end are not counted by ncu"}, f, indent=1)
    print("wrote profiles/%s_*" % tag)
