"""SASS mnemonic counts per kernel of the built library (cuobjdump -sass): the static evidence for the Blackwell
specific instructions (UTMALDG / UTMASTG = TMA tensor load / store, FFMA2 / FADD2 / FMUL2 = packed FP32,
VIMNMX / VIADDMNMX = per-halfword maximum / min(a + b, c), the column pass of the gravity distance transform; FLO / SHF =
the bit scans and funnel shifts of the bitmap kernels) and for what each kernel is made of.

    python scripts/sass_table.py > profiles/r02c_sass.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "detecting_cones_aoslo_b200", "libaoslo_b200.so")
WANT = ['UTMALDG', 'UTMASTG', 'FFMA2', 'FADD2', 'FMUL2', 'VIMNMX', 'VIADDMNMX', 'FLO', 'SHF', 'DFMA', 'DADD', 'DMUL', 'MUFU', 'LDS', 'STS',
        'ATOMG', 'SYNCS', 'LDGSTS', 'BAR']


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r'^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m and cur:
            counts[cur][m.group(2).split('.')[0]] += 1
    names = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for mangled, name in zip(counts, names):
        short = re.sub(r'\(.*', '', name).replace('void ', '').replace('aoslo::', '')
        rows.append((short, sum(counts[mangled].values()), counts[mangled]))
    rows.sort(key=lambda r: -r[1])
    print("# r02c: SASS mnemonics per kernel (`cuobjdump -sass libaoslo_b200.so`, sm_100a; static instruction counts)\n")
    print("| kernel | SASS instr | " + " | ".join(WANT) + " |")
    print("|---|---:|" + "---:|" * len(WANT))
    for short, tot, c in rows:
        if tot >= 40:
            print("| `%s` | %d | " % (short[:70], tot) + " | ".join(str(c[w]) if c[w] else '' for w in WANT) + " |")


if __name__ == "__main__":
    sys.exit(main())
