#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, no GPU needed) into the small tracked CSV that bench.py's `roofline.traffic` reads.

    python tools/ncu_extract.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...] > profiles/r02_ncu_dominant.csv

One row per profiled launch: kernel, grid, duration, DRAM bytes read / written, tensor-pipe activity, L2->SM bytes, L2 hit rate."""
import csv
import subprocess
import sys

WANT = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct", "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_bytes",
        "lts__t_sector_hit_rate.pct": "l2_hit_pct", "sm__cycles_elapsed.avg": "sm_cycles", "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct", "launch__registers_per_thread": "regs", "launch__shared_mem_per_block_dynamic": "smem_dyn"}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}

w = csv.writer(sys.stdout)
cols = ["report", "kernel", "grid"] + list(WANT.values())
w.writerow(cols)
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        rec = {"report": rep.split("/")[-1], "kernel": r[ix["Kernel Name"]][:160], "grid": r[ix["Grid Size"]]}
        for m, name in WANT.items():
            if m in ix:
                v = float(r[ix[m]].replace(",", "") or 0)
                rec[name] = v * UNIT.get(units[ix[m]], 1.0) if name in ("duration_us", "dram_bytes_read", "dram_bytes_write", "l2_to_sm_bytes", "smem_dyn") else v
        w.writerow([rec.get(c, "") for c in cols])
