#!/usr/bin/env python
"""Summarise an .ncu-rep: headline metrics + dynamic instructions per source line (needs -lineinfo)."""
import collections, csv, io, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__cycles_active.avg",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
        "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fmalite.sum", "sm__inst_executed_pipe_fmaheavy.sum"]


def run(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    src_file = sys.argv[2] if len(sys.argv) > 2 else None
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    for row in raw[2:]:
        d = dict(zip(hdr, row))
        print("==", d.get("Kernel Name"), "grid", d.get("Grid Size"), "block", d.get("Block Size"))
        for k in KEYS:
            if k in d:
                print(f"  {k:75s} {d[k]:>16s} {units[hdr.index(k)]}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv", "--print-source", "sass,cuda" if False else "sass"]))))
    # find header row
    hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    h = src[hi]
    iS, iE, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = sum(int(r[iE]) for r in src[hi + 1:] if len(r) > iE and r[iE].isdigit())
    ops = collections.Counter()
    stalls = collections.Counter()
    nsamp = 0
    for r in src[hi + 1:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        t = r[iS].split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += int(r[iE])
        for i in stall_cols:
            if r[i].isdigit():
                stalls[h[i]] += int(r[i])
        nsamp += int(r[iSm]) if r[iSm].isdigit() else 0
    print("total warp instructions", tot)
    print("opcode mix:", ", ".join(f"{o} {100 * c / tot:.1f}%" for o, c in ops.most_common(18)))
    st = sum(stalls.values()) or 1
    print("stall samples:", ", ".join(f"{k[6:]} {100 * v / st:.1f}%" for k, v in stalls.most_common(8)))


if __name__ == "__main__":
    main()
