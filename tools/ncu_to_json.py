""".ncu-rep -> JSON summary rows (one per kernel: the LAST captured launch of each kernel name), the metrics DESIGN.md quotes:
  python tools/ncu_to_json.py gpurun_out/prof.ncu-rep profiles/r02_ncu_kernels.json"""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "sm__cycles_elapsed.avg.per_second"]


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for n, r in enumerate(rows[2:]):
        if len(r) < len(hdr) or "_kernel" not in r[idx["Kernel Name"]] or "at::" in r[idx["Kernel Name"]]:
            continue
        d = {"launch": n, "kernel": r[idx["Kernel Name"]]}
        for k in KEEP:
            if k in idx:
                d[f"{k} [{units[idx[k]]}]"] = r[idx[k]]
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(d["kernel"][:90], "|", d.get("gpu__time_duration.sum [us]", d.get("gpu__time_duration.sum [ms]", "")),
              "| tensor", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active [%]"),
              "| dram R/W", [v for k, v in d.items() if k.startswith("dram__bytes_read.sum [")],
              [v for k, v in d.items() if k.startswith("dram__bytes_write.sum [")],
              "| dram %", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed [%]"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
