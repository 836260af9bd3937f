#!/usr/bin/env python
"""Summarise an `ncu --set full` report: the LAST launch of every kernel, one markdown row each.

    python tools/ncu_summary.py gpurun_out/r2/prof_r2c.ncu-rep [--ids 3,4,5,18] [> profiles/....md]

--ids: report these launch IDs (in this order) instead of the last launch per kernel name.
"""
import csv
import io
import subprocess
import sys

M = {
    "dur": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
    "dram": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l2": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l2b": "l1tex__m_xbar2l1tex_read_bytes.sum", "xbar": "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "l2hit": "lts__t_sector_hit_rate.pct",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor2": "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size",
    "occ": "sm__warps_active.avg.pct_of_peak_sustained_active",
}


def to_f(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return float("nan")


def main():
    rep = sys.argv[1]
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    last = {}
    ids = [a.split("=")[1] if "=" in a else sys.argv[i + 1] for i, a in enumerate(sys.argv) if a.startswith("--ids")]
    if ids:
        by_id = {r[col["ID"]]: r for r in rows[2:]}
        last = {f"#{i} {by_id[i][col['Kernel Name']]}": by_id[i] for i in ids[0].split(",")}
    else:
        for r in rows[2:]:
            last[r[col["Kernel Name"]]] = r        # later launches overwrite the warm-up launch
    pick = lambda r, k: to_f(r[col[M[k]]]) if M[k] in col else float("nan")
    scale = lambda k: {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(units[col[M[k]]].strip(), 1.0) if M[k] in col else 1.0
    bscale = lambda k: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(units[col[M[k]]].strip(), 1.0) if M[k] in col else 1.0
    print("| kernel | duration us | DRAM read / write GB | DRAM GB/s | DRAM % of peak | L2 -> SM GB (% of xbar peak) | L2 throughput % | L2 hit % | tensor pipe % | issue slots % | warps active % | regs | grid x block |")
    print("|---|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
    for name, r in last.items():
        dur = pick(r, "dur") * scale("dur")
        rd, wr = pick(r, "rd") * bscale("rd"), pick(r, "wr") * bscale("wr")
        l2b = pick(r, "l2b") * bscale("l2b")
        tp = pick(r, "tensor")
        if tp != tp:
            tp = pick(r, "tensor2")
        short = name if len(name) < 90 else name[:87] + "..."
        print(f"| `{short}` | {dur:.1f} | {rd / 1e9:.3f} / {wr / 1e9:.3f} | {(rd + wr) / dur / 1e3:.0f} | {pick(r, 'dram'):.1f} | {l2b / 1e9:.2f} ({pick(r, 'xbar'):.1f}) | "
              f"{pick(r, 'l2'):.1f} | {pick(r, 'l2hit'):.1f} | {tp:.1f} | {pick(r, 'issue'):.1f} | {pick(r, 'occ'):.1f} | {int(pick(r, 'regs'))} | "
              f"{int(pick(r, 'grid'))} x {int(pick(r, 'block'))} |")


if __name__ == "__main__":
    main()
