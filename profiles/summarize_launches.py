"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki]
        m = re.search(r"(\w+_kernel|\w+Kernel\w*|at::native::\w+)", name)
        name = m.group(1) if m else name[:60]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':44s} {'launches':>8s} {'total_ms':>10s} {'share':>7s} {'avg_us':>9s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k[:44]:44s} {v[0]:8d} {v[1] / 1e3:10.3f} {v[1] / tot * 100:6.1f}% {v[1] / v[0]:9.1f}")
    print(f"{'TOTAL':44s} {sum(v[0] for v in agg.values()):8d} {tot / 1e3:10.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
