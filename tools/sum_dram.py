"""Sum dram bytes / kernel time over an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`
log (one decode), per kernel and in total:  python tools/sum_dram.py log.csv [algorithmic_GB]"""
import csv
import collections
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10]
hdr = rows[0]
i_name, i_metric, i_unit, i_val = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
i_id = hdr.index("ID")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}
per = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
seen = set()
for r in rows[1:]:
    name = r[i_name].split("(")[0].replace("void ", "").replace("dk::", "")
    v = float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
    e = per[name]
    if (r[i_id], name) not in seen:
        seen.add((r[i_id], name))
        e[0] += 1
    if r[i_metric] == "dram__bytes_read.sum":
        e[1] += v
    elif r[i_metric] == "dram__bytes_write.sum":
        e[2] += v
    elif r[i_metric] == "gpu__time_duration.sum":
        e[3] += v
tot = [sum(e[k] for e in per.values()) for k in range(4)]
print(f"{'kernel':58s} {'n':>4s} {'read GB':>8s} {'write GB':>8s} {'ms':>8s} {'GB/s':>7s}")
for name, e in sorted(per.items(), key=lambda kv: -kv[1][3]):
    print(f"{name[:58]:58s} {e[0]:4d} {e[1] / 1e9:8.3f} {e[2] / 1e9:8.3f} {e[3] * 1e3:8.3f} {(e[1] + e[2]) / max(e[3], 1e-12) / 1e9:7.0f}")
print(f"{'TOTAL':58s} {tot[0]:4d} {tot[1] / 1e9:8.3f} {tot[2] / 1e9:8.3f} {tot[3] * 1e3:8.3f} {(tot[1] + tot[2]) / max(tot[3], 1e-12) / 1e9:7.0f}")
if len(sys.argv) > 2:
    alg = float(sys.argv[2])
    print(f"traffic / algorithmic ({alg} GB): {(tot[1] + tot[2]) / 1e9 / alg:.3f}")
