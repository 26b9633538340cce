"""Per-kernel table of one detection run from an `ncu --metrics gpu__time_duration.sum --csv` launch list.

    python scripts/launch_table.py gpurun_out/launches.csv [run_index]

(Conditional graph nodes cannot be profiled per kernel, so the list is taken with AOSLO_NO_GRAPH=1: the same
kernels launched eagerly, loop conditions evaluated on the host.)"""
import collections
import csv
import re
import sys


def table(path, run=1):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    names = [(r['Kernel Name'], float(r['Metric Value'].replace(',', ''))) for r in rows]
    idx = [i for i, (n, _) in enumerate(names) if 'blobness' in n]
    start, end = idx[run], idx[run + 1]
    agg = collections.OrderedDict()
    for n, v in names[start:end]:
        key = re.sub(r'\(.*', '', n)
        a = agg.setdefault(key, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += v
        a[2] = max(a[2], v)
    tot = sum(a[1] for a in agg.values())
    out = ["launches in the run: %d, summed kernel time: %.1f us" % (end - start, tot / 1e3), "",
           "| kernel | launches | total us | share | avg us | max us |", "|---|---:|---:|---:|---:|---:|"]
    for k, (c, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.1f | %.1f%% | %.2f | %.2f |" % (k.replace('void ', '').replace('aoslo::', '')[:70],
                                                                    c, t / 1e3, 100 * t / tot, t / 1e3 / c, m / 1e3))
    return "\n".join(out)


if __name__ == "__main__":
    print(table(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1))
