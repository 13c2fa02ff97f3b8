#!/usr/bin/env python
"""Reads an `ncu --set full` report (here, no GPU needed) and writes the per-launch DRAM traffic of the kernels whose name
matches a pattern to a small JSON that bench.py quotes as `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum).

    python tools/ncu_traffic.py gpurun_out/r02b_spmm.ncu-rep spmm_fan profiles/r02_spmm_traffic.json "tools/run_kernels.py spmm"
"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "second": 1e3, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}


def main():
    rep, pattern, out, cmd = sys.argv[1:5]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    num = lambda r, k: float(r[col[k]].replace(",", ""))
    launches = []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pattern not in name:
            continue
        rd = num(r, "dram__bytes_read.sum") * UNIT[units[col["dram__bytes_read.sum"]]]
        wr = num(r, "dram__bytes_write.sum") * UNIT[units[col["dram__bytes_write.sum"]]]
        launches.append({"kernel": name.split("(")[0].replace("void <unnamed>::", ""), "grid": int(num(r, "launch__grid_size")),
                         "block": int(num(r, "launch__block_size")), "registers": int(num(r, "launch__registers_per_thread")),
                         "ms_under_ncu": num(r, "gpu__time_duration.sum") * TIME[units[col["gpu__time_duration.sum"]]],
                         "dram_read_bytes": rd, "dram_write_bytes": wr,
                         "dram_pct_of_peak": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                         "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct"),
                         "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active")})
    json.dump({"source": rep, "command": cmd, "how": "ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum per launch",
               "launches": launches}, open(out, "w"), indent=1)
    for l in launches:
        print(l["kernel"], l["grid"], f"{l['ms_under_ncu']:.3f} ms", f"{(l['dram_read_bytes'] + l['dram_write_bytes']) / 1e9:.2f} GB")


if __name__ == "__main__":
    main()
