"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: share of the step per kernel."""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in r:
        if len(row) < len(hdr):
            continue
        val = float(row[idx["Metric Value"]].replace(",", ""))
        unit = row[idx["Metric Unit"]]
        val *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        name = re.sub(r"void |dk::", "", re.sub(r"\(.*", "", row[idx["Kernel Name"]]))
        agg[name][0] += 1
        agg[name][1] += val
        tot += val
    print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms summed kernel time "
          f"(ncu serialises launches, cold caches: compare SHARES)")
    print(f"{'share':>7s} {'launches':>8s} {'total ms':>10s} {'avg us':>9s}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / tot * 100:6.2f}% {v[0]:8d} {v[1] / 1e3:10.3f} {v[1] / v[0]:9.1f}  {k[:110]}")


if __name__ == "__main__":
    main(sys.argv[1])
