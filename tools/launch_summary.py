"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X ...`): per-kernel launches, total and
share of device time; with --per-step N divides by N forward passes; with --steps-by KERNEL the window is trimmed to the whole steps
between the first and the last launch of KERNEL (e.g. validate_inputs_kernel, the first launch of a forward) and divided by their number.
   python tools/launch_summary.py gpurun_out/launches.csv [--per-step N | --steps-by validate_inputs]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    per = int(sys.argv[sys.argv.index("--per-step") + 1]) if "--per-step" in sys.argv else 1
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if r[mi] != "gpu__time_duration.sum":
            continue
        v = float(r[vi].replace(",", ""))
        v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[ui], 1.0)
        rows.append((re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip(), v))
    if "--steps-by" in sys.argv:
        mk = sys.argv[sys.argv.index("--steps-by") + 1]
        idx = [i for i, (k, _) in enumerate(rows) if mk in k]
        if len(idx) >= 2:
            rows = rows[idx[0]:idx[-1]]
            per = len(idx) - 1
    agg = OrderedDict()
    for k, v in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v in rows)
    print("launches %d (%.1f per step), device time %.1f us (%.1f us per step)" % (len(rows), len(rows) / per, tot, tot / per))
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%6.1f%%  %8.1f us/step  %5.1f launches/step  %6.1f us avg  %s" % (100 * v / tot, v / per, n / per, v / n, k))


if __name__ == "__main__":
    main()
