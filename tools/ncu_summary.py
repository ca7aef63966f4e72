#!/usr/bin/env python
"""Turn an `ncu --set full` report (read with the ncu CLI present in this image) into the small tracked artefacts:

    python tools/ncu_summary.py gpurun_out/<name>.ncu-rep profiles/<tag>        # -> <tag>_summary.csv (+ prints the table)
    python tools/ncu_summary.py gpurun_out/<name>.ncu-rep profiles/<tag> --traffic gs_mesh_1M_1080p   # also updates profiles/ncu_traffic.json

The summary keeps, per captured launch: duration, warp instructions, IPC, issue-active, pipe utilisation, occupancy, registers,
DRAM bytes read/written and the top stall reasons.  `ncu_traffic.json` is what bench.py reads `roofline.traffic` from."""
import csv
import json
import os
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum"]
SLOT = {"k_composite_bwd": "composite_bwd", "k_composite_fwd": "composite_fwd", "k_preprocess_bwd": "preprocess_bwd",
        "k_preprocess_fwd": "preprocess_fwd", "k_adam_sh": "adam_sh", "k_adam": "adam", "k_bin_tiles": "bin_tiles"}


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    out = [["kernel"] + KEEP + ["top_stalls (warps per issue-active cycle)"]]
    traffic = {}
    for r in data:
        name = r[col["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "")
        vals = [r[col[k]] if k in col else "" for k in KEEP]
        st = sorted(((float(r[col[h]] or 0), h) for h in stall), reverse=True)[:4]
        tops = "; ".join(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, h in st)
        out.append([short] + vals + [tops])
        u = {k: units[col[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in col}
        mult = lambda unit: {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        for pre, slot in SLOT.items():
            if short.startswith(pre) and slot not in traffic:
                traffic[slot] = {"dram_bytes_read": float(r[col["dram__bytes_read.sum"]]) * mult(u["dram__bytes_read.sum"]),
                                 "dram_bytes_write": float(r[col["dram__bytes_write.sum"]]) * mult(u["dram__bytes_write.sum"]),
                                 "kernel": short, "capture": os.path.basename(tag) + "_summary.csv"}
                break
    with open(tag + "_summary.csv", "w", newline="") as f:
        csv.writer(f).writerows(out)
    for row in out[1:]:
        print(row[0][:40], "|", " ".join(f"{k.split('.')[0].split('__')[-1]}={v[:9]}" for k, v in zip(KEEP[:8], row[1:9])), "|", row[-1])
    if "--traffic" in sys.argv:
        wl = sys.argv[sys.argv.index("--traffic") + 1]
        p = os.path.join(os.path.dirname(tag) or ".", "ncu_traffic.json")
        d = json.load(open(p)) if os.path.exists(p) else {}
        d.setdefault(wl, {}).update(traffic)
        json.dump(d, open(p, "w"), indent=1, sort_keys=True)
        print("updated", p)


if __name__ == "__main__":
    main()
