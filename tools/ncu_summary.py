#!/usr/bin/env python
"""Turns an `ncu --set full` report into the two files bench.py and DESIGN.md cite:

  profiles/ncu_<tag>_summary.md   one table row per launch (time, DRAM bytes, pipe utilisation, occupancy)
  profiles/ncu_<tag>_traffic.json per-kernel and per-stage DRAM bytes per launch (roofline.traffic in bench.py)

usage: python tools/ncu_summary.py gpurun_out/full_r1c.ncu-rep r1c [--seqs 64] [--command "..."]
Needs the `ncu` CLI (reads the report here, no GPU needed).
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

COLS = [
    ("time ms", "gpu__time_duration.sum"),
    ("dram rd MB", "dram__bytes_read.sum"),
    ("dram wr MB", "dram__bytes_write.sum"),
    ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("sm %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps act %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("issue %", "sm__inst_issued.avg.pct_of_peak_sustained_active"),
    ("warp inst", "smsp__inst_executed.sum"),
    ("regs", "launch__registers_per_thread"),
    ("grid", "launch__grid_size"),
    ("block", "launch__block_size"),
    ("L1 hit %", "l1tex__t_sector_hit_rate.pct"),
    ("L2 hit %", "lts__t_sector_hit_rate.pct"),
    ("fmaheavy %", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("fma %", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("alu %", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    ("lsu %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
]

# stage -> kernels (name prefix[:count per step]) as bench.py's stage timers group them
STAGES = {
    "image_pyramid": ["k_blur3", "k_pyrdown:3"],
    "lm_solve": ["k_lm_solve", "k_track_pose"],
    "blur_select": ["k_blur3", "k_fill_planes", "k_select"],
    "disparity": ["k_disparity<1", "k_disparity<0"],
    "depth_refine": ["k_raw8", "k_refine"],
    "depth_pyr_promote_prepare": ["k_track_finalize", "k_lm_count", "k_lm_scan", "k_lm_fill"],
}

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
              "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def short_name(full: str) -> str:
    n = re.sub(r"<unnamed>::|void |\(anonymous namespace\)::", "", full)
    return n.split("(")[0].strip()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("tag")
    ap.add_argument("--seqs", type=int, default=64)
    ap.add_argument("--command", default="")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def val(r, metric):
        if metric not in ix:
            return None
        s = r[ix[metric]].replace(",", "")
        try:
            v = float(s)
        except ValueError:
            return None
        return v * UNIT_SCALE.get(units[ix[metric]], 1.0)

    lines = [f"# ncu --set full, {args.tag} ({args.seqs} KITTI frames per launch)", ""]
    if args.command:
        lines += [f"Command: `{args.command}`", ""]
    lines += ["Per-launch times under ncu are cold-cache and serialised: compare SHARES with the bench's stage timers, not "
              "absolutes. DRAM bytes are `dram__bytes_read.sum` / `dram__bytes_write.sum` of that launch.", ""]
    lines.append("| kernel | " + " | ".join(c[0] for c in COLS) + " |")
    lines.append("|---|" + "---|" * len(COLS))
    per_kernel = {}
    for r in body:
        name = short_name(r[ix["Kernel Name"]])
        cells = []
        for label, metric in COLS:
            v = val(r, metric)
            if v is None:
                cells.append("NA")
            elif label in ("dram rd MB", "dram wr MB"):
                cells.append(f"{v / 1e6:.2f}")
            elif label == "time ms":
                cells.append(f"{v:.4f}")
            elif label in ("warp inst",):
                cells.append(f"{v:.3e}")
            elif label in ("regs", "grid", "block"):
                cells.append(f"{int(v)}")
            else:
                cells.append(f"{v:.1f}")
        lines.append(f"| {name} | " + " | ".join(cells) + " |")
        k = per_kernel.setdefault(name, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0, "warp_inst": 0.0, "issue_pct": 0.0})
        k["launches"] += 1
        k["warp_inst"] += val(r, "smsp__inst_executed.sum") or 0.0
        k["issue_pct"] += val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active") or 0.0
        k["dram_bytes"] += (val(r, "dram__bytes_read.sum") or 0.0) + (val(r, "dram__bytes_write.sum") or 0.0)
        k["ms"] += val(r, "gpu__time_duration.sum") or 0.0
    kernels = {n: {"dram_bytes_per_launch": k["dram_bytes"] / k["launches"], "ms_per_launch": k["ms"] / k["launches"],
                   "warp_inst_per_launch": k["warp_inst"] / k["launches"], "issue_active_pct": k["issue_pct"] / k["launches"],
                   "launches_in_capture": k["launches"]} for n, k in per_kernel.items()}
    stage_bytes, stage_ms = {}, {}
    for stage, pats in STAGES.items():
        tb, tm, found = 0.0, 0.0, False
        for pat in pats:
            prefix, _, cnt = pat.partition(":")
            for n, k in kernels.items():
                if n.startswith(prefix):
                    tb += k["dram_bytes_per_launch"] * (int(cnt) if cnt else 1)
                    tm += k["ms_per_launch"] * (int(cnt) if cnt else 1)
                    found = True
        if found:
            stage_bytes[stage], stage_ms[stage] = tb, tm
    tot = sum(stage_ms.values())
    lines += ["", "## Stage shares (serialised ncu times)", "", "| stage | ms | share | DRAM MB |", "|---|---|---|---|"]
    for s in stage_ms:
        lines.append(f"| {s} | {stage_ms[s]:.4f} | {100 * stage_ms[s] / tot:.1f} % | {stage_bytes[s] / 1e6:.1f} |")
    out_md = os.path.join(ROOT, "profiles", f"ncu_{args.tag}_summary.md")
    open(out_md, "w").write("\n".join(lines) + "\n")
    out_js = os.path.join(ROOT, "profiles", f"ncu_{args.tag}_traffic.json")
    json.dump({"source": f"profiles/ncu_{args.tag}_summary.md (ncu --set full, {args.seqs} KITTI frames per launch)",
               "sequences_per_launch": args.seqs, "kernels": kernels, "stages": STAGES,
               "stage_dram_bytes": stage_bytes, "stage_ms_serialised": stage_ms}, open(out_js, "w"), indent=1)
    print(out_md, out_js)


if __name__ == "__main__":
    main()
