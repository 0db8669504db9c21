"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per-kernel count, total, average, share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 2:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0][:70]
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e6:.3f} ms device time (serialised, cold cache)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} n={v[0]:5d} total={v[1] / 1e6:9.3f} ms avg={v[1] / v[0] / 1e3:9.1f} us share={v[1] / tot:.3f}")


if __name__ == "__main__":
    main(sys.argv[1])
