"""Extract per-kernel facts from an `ncu --set full` report into profiles/r<round>_ncu_kernels.json.
usage: ncu -i gpurun_out/prof.ncu-rep --page raw --csv | python profiles/ncu_kernels.py profiles/r<round>_ncu_kernels.json"""
import csv
import json
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}


def val(r, k):
    if k not in col:
        return None
    try:
        v = float(r[col[k]].replace(",", ""))
    except ValueError:
        return None
    u = units[col[k]]
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(u, 1)


out = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]].split("(")[0].replace("<unnamed>::", "")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    out[name] = {
        "grid": r[col["launch__grid_size"]], "block": r[col["launch__block_size"]],
        "duration_s_under_ncu": val(r, "gpu__time_duration.sum"),
        "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes": (rd or 0) + (wr or 0),
        "dram_throughput_pct_of_peak": val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "warp_instructions": val(r, "smsp__inst_executed.sum"),
        "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": val(r, "launch__registers_per_thread"),
        "fma_pipe_active_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    }
json.dump(out, open(sys.argv[1], "w"), indent=1)
print(json.dumps(out, indent=1))
