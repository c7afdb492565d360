"""Summarise an `ncu --set full` report into JSON: per captured launch the kernel name, grid/block, duration, DRAM
bytes read + written, DRAM / tensor-pipe / SM utilisation, registers, and the top warp-stall reasons.

    ncu -i gpurun_out/x.ncu-rep --page raw --csv > x.csv ; python tools/ncu_summarise.py x.csv > profiles/x.json
(or pass the .ncu-rep directly: the script calls ncu itself)."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__cycles_active.avg": "smsp_cycles_active",
}


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return v * mult


def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6, "second": 1e6}.get(unit, 1)


def main():
    path = sys.argv[1]
    if path.endswith(".ncu-rep"):
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    else:
        text = open(path).read()
    rows = list(csv.reader(io.StringIO(text)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    head, units = rows[hi], rows[hi + 1]
    out = []
    for r in rows[hi + 2:]:
        if len(r) != len(head):
            continue
        rec = {"kernel": r[head.index("Kernel Name")][:90]}
        stalls = []
        for j, name in enumerate(head):
            if name in WANT:
                try:
                    v = float(r[j].replace(",", ""))
                except ValueError:
                    continue
                key = WANT[name]
                if key in ("dram_read", "dram_write"):
                    v = to_bytes(v, units[j])
                elif key == "duration":
                    v = to_us(v, units[j])
                    key = "duration_us"
                rec[key] = v
            elif name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
                try:
                    stalls.append((float(r[j].replace(",", "")), name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        if "dram_read" in rec and "dram_write" in rec:
            rec["dram_bytes_per_launch"] = rec["dram_read"] + rec["dram_write"]
            if rec.get("duration_us"):
                rec["dram_GBps_under_ncu"] = round(rec["dram_bytes_per_launch"] / rec["duration_us"] / 1e3, 1)
        rec["top_stalls"] = [[n, round(v, 2)] for v, n in sorted(stalls, reverse=True)[:4]]
        out.append(rec)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
