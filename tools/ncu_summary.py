#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): per-kernel roofline counters and the source
lines that collect the most warp-stall samples.

    python tools/ncu_summary.py gpurun_out/prof_fv_r01.ncu-rep [--kernel regex] [--top 25] > profiles/...

Uses `ncu -i <rep> --page raw --csv` and `--page source --csv --print-source cuda,sass`
(the library is compiled with -lineinfo).
"""
import argparse, csv, io, re, subprocess, sys

RAW = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thr"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/tex %peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp inst"),
    ("sm__sass_inst_executed_op_shared_ld.sum", "LDS inst"), ("sm__sass_inst_executed_op_shared_st.sum", "STS inst"),
    ("sm__sass_inst_executed_op_global_ld.sum", "LDG inst"), ("sm__sass_inst_executed_op_global_st.sum", "STG inst"),
]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def raw_table(rep):
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv", "--print-kernel-base", "demangled"]))))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        rec = dict(zip(hdr, r))
        print(f"== {rec.get('Kernel Name', '?')}   [id {rec.get('ID', '?')}]")
        for key, label in RAW:
            if key in rec and rec[key] != "":
                print(f"   {label:<26} {rec[key]:>16} {units[hdr.index(key)]}")
        rd, wr = rec.get("dram__bytes_read.sum"), rec.get("dram__bytes_write.sum")
        print()


def source_table(rep, kernel, top):
    out = run([rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name-base", "demangled",
               "-k", f"regex:{kernel}"])
    rows = list(csv.reader(io.StringIO(out)))
    fname, func, hdr = None, None, None
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            func = r[1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr and r[0].isdigit():
            try:
                lines.append((int(r[4] or 0), int(r[5] or 0), fname, int(r[0]), r[1].strip(), func))
            except ValueError:
                pass
    by_func = {}
    for l in lines:
        by_func.setdefault(l[5], []).append(l)
    for func, ls in by_func.items():
        tot = sum(l[0] for l in ls) or 1
        print(f"== stall samples by source line: {func}  (total {tot})")
        for s, ni, f, ln, src, _ in sorted(ls, reverse=True)[:top]:
            if s == 0:
                break
            print(f"   {100.0 * s / tot:5.1f}%  {s:7d}  {f}:{ln:<4d} {src[:110]}")
        print()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", default=".")
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args()
    raw_table(a.rep)
    source_table(a.rep, a.kernel, a.top)
