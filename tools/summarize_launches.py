"""Summarise an `ncu --metrics gpu__time_duration.sum,... --csv` launch list: per kernel, over the LAST forward.

    python tools/summarize_launches.py gpurun_out/launches_b16.csv
"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            hdr, start = r, i + 1
            break
    ki, vi, mi, ii = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("ID")
    L = collections.OrderedDict()
    for r in rows[start:]:
        if len(r) <= vi:
            continue
        d = L.setdefault(int(r[ii]), {"name": re.sub(r"\(.*", "", r[ki]).replace("catseg::", "").replace("void ", "")})
        try:
            d[r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            pass
    ids = sorted(L)
    starts = [i for i in ids if "normalize_img" in L[i]["name"] or "inv_norm_pixels" in L[i]["name"]]
    last, end = starts[-1], ids[-1] + 1
    if len(starts) > 1 and end - last < last - starts[-2]:      # the capture stopped inside the last forward: use the one before
        last, end = starts[-2], starts[-1]
    agg = collections.OrderedDict()
    for i in ids:
        if i < last or i >= end:
            continue
        d = L[i]
        a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0)
        a[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        a[3] += d.get("smsp__inst_executed.sum", 0)
        a[4] += d.get("sm__cycles_elapsed.max", 0)
        a[5] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0) * d.get("gpu__time_duration.sum", 0)
    tot = sum(a[1] for a in agg.values())
    print(f"last forward: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms (ncu: cold cache, serialised)")
    print(f"{'kernel':58s} {'n':>3s} {'ms':>8s} {'share':>6s} {'DRAM GB':>8s} {'GB/s':>6s} {'IPC/SM':>6s} {'tensor%':>7s}")
    for n, a in agg.items():
        ipc = a[3] / 148 / max(a[4], 1)
        print(f"{n[:58]:58s} {a[0]:3d} {a[1] / 1e6:8.3f} {100 * a[1] / tot:5.1f}% {a[2] / 1e9:8.2f} {a[2] / max(a[1], 1):6.0f} {ipc:6.2f} "
              f"{a[5] / max(a[1], 1):7.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
