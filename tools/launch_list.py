#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> the per-kernel summary kept under profiles/.

    python tools/launch_list.py gpurun_out/launches.csv "header line" [steps] > profiles/rNN_launch_list_....txt
"""
import collections
import csv
import sys


def main():
    path, header = sys.argv[1], sys.argv[2]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    rows = list(csv.reader(ln for ln in open(path) if ln.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v * (1000.0 if r[ui] in ("ms", "msecond") else 1.0)
        t, n = agg.get(r[ki], (0.0, 0))
        agg[r[ki]] = (t + v, n + 1)
    total = sum(t for t, _ in agg.values())
    print(f"# {header}")
    print(f"# total {total:.1f} us over {sum(n for _, n in agg.values())} launches = {steps} step(s); per-launch times under ncu are serialised and cold-cache: compare SHARES")
    for k, (t, n) in agg.items():
        print(f"{t:10.1f} us  {100.0 * t / total:5.1f} %  x{n:<3d} {k[:150]}")


if __name__ == "__main__":
    main()
