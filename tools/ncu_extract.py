#!/usr/bin/env python
"""Turn an ``.ncu-rep`` (ncu --set full) into the small ``metric,unit,value`` CSVs kept under profiles/.

    python tools/ncu_extract.py gpurun_out/prof.ncu-rep pairwise_tile profiles/r1_ncu_full_pairwise.csv [--largest]

Picks the first launch whose kernel name contains the substring (``--largest``: the one with the longest
gpu__time_duration), transposes ncu's raw page and keeps the metrics that the roofline discussion uses."""
import csv
import io
import subprocess
import sys

KEEP = ("dram__bytes", "dram__throughput", "gpu__time_duration", "sm__throughput", "sm__pipe", "sm__inst_executed_pipe",
        "sm__warps_active", "sm__cycles_active.avg", "smsp__issue_active", "smsp__inst_executed.sum", "l1tex__data_bank",
        "l1tex__data_pipe_lsu_wavefronts", "lts__t_sector_hit_rate", "lts__throughput", "lts__t_bytes.sum",
        "launch__registers", "launch__occupancy", "launch__shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__average_warp", "smsp__warp_issue_stalled", "smsp__average_warps_issue_stalled", "sm__mem_tensor",
        "l1tex__m_xbar2l1tex_read_bytes", "smsp__cycles_active.avg", "sm__inst_executed.sum", "smsp__pcsamp_warps_issue")


def main():
    rep, pattern, out = sys.argv[1:4]
    largest = "--largest" in sys.argv
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    names, units, launches = rows[0], rows[1], rows[2:]
    kcol, dcol = names.index("Kernel Name"), None
    for i, nm in enumerate(names):
        if nm.endswith("gpu__time_duration.sum"):
            dcol = i
    match = [r for r in launches if pattern in r[kcol]]
    if not match:
        raise SystemExit(f"no launch matches {pattern!r}: {sorted(set(r[kcol][:60] for r in launches))}")
    row = max(match, key=lambda r: float(r[dcol] or 0)) if largest and dcol is not None else match[0]
    with open(out, "w") as f:
        f.write("metric,unit,value\n")
        for key in ("Kernel Name", "Block Size", "Grid Size"):
            f.write(f'"{key}","","{row[names.index(key)]}"\n')
        for nm, unit, val in zip(names, units, row):
            base = nm.split(".", 2)[-1] if nm.split(".")[0].isupper() or "Triage" in nm else nm
            if val != "" and any(k in base for k in KEEP):
                f.write(f'"{base}","{unit}","{val}"\n')
    print(f"{out}: {row[kcol][:80]} ({len(match)} matching launches)")


if __name__ == "__main__":
    main()
