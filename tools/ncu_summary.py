#!/usr/bin/env python
"""Condense an Nsight Compute report (.ncu-rep, read here on the CPU box with `ncu -i`) into the small CSV the
repo keeps under profiles/: one row per profiled launch with duration, DRAM traffic, tensor-pipe activity,
achieved occupancy inputs and instruction counts.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_ncu_cfg3.csv
"""
import csv
import subprocess
import sys

KEEP = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "duration_us"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clock_ghz"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct_active"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pct_elapsed"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "ipc_per_sm"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(k), name) for k, name in KEEP if k in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([name + ("" if not units[i] or name == "kernel" else " [%s]" % units[i]) for i, name in idx])
        for r in rows[2:]:
            vals = []
            for i, name in idx:
                v = r[i]
                if name == "kernel":
                    v = v.replace("qnnb::<unnamed>::", "").replace("void ", "")[:110]
                vals.append(v)
            w.writerow(vals)
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
