"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (launches, total us, share)."""
import collections
import csv
import sys


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    unit = rows[1][hdr.index("Metric Unit")] if "Metric Unit" in hdr else "ns"
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit.strip(), 1e-3)
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + float(r[iv].replace(",", "")) * scale)
    total = sum(t for _, t in agg.values())
    print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'share':>7s}")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:70]:70s} {n:8d} {t:12.1f} {100 * t / total:6.1f}%")
    print(f"{'TOTAL':70s} {sum(n for n, _ in agg.values()):8d} {total:12.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
