#!/usr/bin/env python
"""Summarise an `ncu --set full` report for profiles/: one row per captured kernel launch.

    python scripts/ncu_summary.py gpurun_out/r01f/prof.ncu-rep profiles/r01f_ncu_full_T1279.json

Reads the report with `ncu -i ... --page raw --csv` (works without a GPU) and keeps the metrics DESIGN.md and
bench.py quote: duration, DRAM bytes read / written, registers, achieved warps, fp64 pipe utilisation, issue
utilisation and the stall-sample breakdown."""
import csv
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "launch__grid_size": "grid",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_elapsed.avg": "sm_cycles",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
         "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}


def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")].replace("void <unnamed>::", "").split("(")[0]}
        stalls = {}
        for i, h in enumerate(hdr):
            v = r[i].replace(",", "")
            if h in KEEP:
                try:
                    x = float(v)
                except ValueError:
                    continue
                if units[i] in SCALE and KEEP[h] in ("time", "dram_read", "dram_write", "dyn_smem"):
                    x *= SCALE[units[i]]
                d[KEEP[h]] = x
            elif h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls[h[len("smsp__pcsamp_warps_issue_stalled_"):]] = int(float(v))
                except ValueError:
                    pass
        if "time" in d:
            d["time_ms"] = d.pop("time") * 1e3
            d["dram_read_GB"] = d.pop("dram_read", 0.0) / 1e9
            d["dram_write_GB"] = d.pop("dram_write", 0.0) / 1e9
            d["dram_GBps"] = (d["dram_read_GB"] + d["dram_write_GB"]) / (d["time_ms"] * 1e-3)
            if d.get("warp_instructions") and d.get("sm_cycles"):
                d["issue_slot_util"] = d["warp_instructions"] / (d["sm_cycles"] * 148 * 4)
        tot = sum(stalls.values()) or 1
        d["stall_samples_pct"] = {k: round(100.0 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6]}
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    print("| kernel | time ms | dram read GB | dram write GB | DRAM GB/s | regs | warps active % | fp64 pipe % | issue util | top stalls |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for d in res:
        st = ", ".join(f"{k} {v}" for k, v in list(d["stall_samples_pct"].items())[:3])
        print(f"| {d['kernel']} | {d['time_ms']:.3f} | {d['dram_read_GB']:.3f} | {d['dram_write_GB']:.3f} | {d['dram_GBps']:.0f} | "
              f"{int(d.get('regs', 0))} | {d.get('warps_active_pct', 0):.1f} | {d.get('fp64_pipe_pct', 0):.1f} | "
              f"{d.get('issue_slot_util', 0):.2f} | {st} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
