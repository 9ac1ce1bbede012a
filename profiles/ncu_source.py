"""Per-instruction hot spots of one kernel from an .ncu-rep captured with --import-source on.

    python profiles/ncu_source.py gpurun_out/prof.ncu-rep [min_pct]
"""
import collections
import csv
import subprocess
import sys


def main(path, min_pct=1.5):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    data, agg = [], {k: 0 for k in keys}
    for r in rows[2:]:
        try:
            n = int(r[ci["Instructions Executed"]])
            s = int(r[ci["# Samples"]])
        except (ValueError, IndexError):
            continue
        for k in keys:
            agg[k] += int(r[ci[k]] or 0)
        data.append((n, s, r[ci["Source"]].strip(), {k[6:]: int(r[ci[k]] or 0) for k in keys if int(r[ci[k]] or 0) > 5}))
    tot = sum(d[0] for d in data)
    stot = sum(d[1] for d in data)
    print("total warp inst", tot, "samples", stot)
    print("stalls:", [(k[6:], v, f"{100 * v / max(stot, 1):.1f}%") for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]])
    print("common exec counts", collections.Counter(d[0] for d in data).most_common(6))
    for i, (n, s, src, st) in enumerate(data):
        if s > stot * min_pct / 100:
            print(f"{i:4d} {n:8d} smp {100 * s / stot:5.1f}%  {src[:70]:70s}", st)


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.5)
