"""Warp-stall sampling summary of one kernel from an `ncu --set full --import-source on` report (SASS view):
  python tools/ncu_stalls.py gpurun_out/prof.ncu-rep regex:attention_fwd_v3 [launch-skip] [top-N]"""
import csv
import subprocess
import sys


def main(rep, kern, skip="0", topn="40"):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--launch-skip", skip,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    print(rows[0][1][:120])
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) or r[idx["# Samples"]] == "# Samples":
            if len(r) > 1 and r[1] == "Source":
                break   # second view (CUDA-C): stop
            continue
        data.append(r)
    n = lambda r, k: int(r[idx[k]] or 0)
    tot = sum(n(r, "# Samples") for r in data)
    print("SASS lines", len(data), "samples", tot)
    agg = {s: sum(n(r, s) for r in data) for s in stalls}
    for s, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
        print(f"  {s:26s} {v:8d} {100 * v / max(tot, 1):5.1f}%")
    print("top lines (samples, executed, SASS, top stalls)")
    order = sorted(range(len(data)), key=lambda i: -n(data[i], "# Samples"))[: int(topn)]
    for i in sorted(order):
        r = data[i]
        st = sorted(((s, n(r, s)) for s in stalls if n(r, s) > 0), key=lambda x: -x[1])[:3]
        print(f"{i:5d} {n(r, '# Samples'):6d} {n(r, 'Instructions Executed'):9d}  {r[idx['Source']].strip()[:64]:64s} {st}")


if __name__ == "__main__":
    main(*sys.argv[1:])
