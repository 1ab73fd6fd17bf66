#!/usr/bin/env python
"""Turn the ncu outputs a gpurun call brought back into the small, tracked
summaries under profiles/ (gpurun_out/ itself is scratch and git-ignored).

    python profiles/summarize.py launches gpurun_out/X_launches.csv profiles/X_launches.md "<command line>"
    python profiles/summarize.py full     gpurun_out/X_prof.ncu-rep profiles/X_full.md     "<command line>"

`launches`: the `--metrics gpu__time_duration.sum` pass -> per-kernel launch count,
total device time and SHARE of the profiled region (cold-cache, serialised times:
only the shares are meaningful).
`full`: one `--set full` capture -> per-launch duration, DRAM bytes, pipe/issue
utilisation, occupancy, registers, plus the top stall reasons.
"""
import collections
import csv
import io
import subprocess
import sys


def launches(src, dst, cmd):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO("".join(lines))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("nsagp::", "")
        key = (name, row["Block Size"], row["Grid Size"])
        v = float(row["Metric Value"].replace(",", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Command: `%s`\n\nSource: `%s`. Times are cold-cache and serialised by the profiler: read the SHARE column.\n\n" % (cmd, src))
        f.write("| kernel | block | grid | launches | total ms | share |\n|---|---|---|---:|---:|---:|\n")
        for (name, blk, grd), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %s | %s | %d | %.3f | %.2f%% |\n" % (name, blk, grd, n, t / 1e6, 100 * t / tot))
        f.write("\nTotal profiled device time: %.1f ms over %d launches.\n" % (tot / 1e6, sum(a[0] for a in agg.values())))


FULL_METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe % (active SMs)"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 inst % (active SMs)"),
    ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "FP64 tensor (DMMA) pipe % (active SMs)"),
    ("sm__ops_path_tensor_src_fp64.sum.per_second", "FP64 tensor ops/ns, chip (peak 37170)"),
    ("sm__ops_path_tensor_src_fp64.sum.pct_of_peak_sustained_elapsed", "FP64 tensor ops % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 sector hit rate"),
    ("lts__t_sector_hit_rate.pct", "L2 sector hit rate"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__cycles_elapsed.max", "cycles"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slot busy %"),
]


def full(src, dst, cmd):
    # src: a .ncu-rep, or the CSV `ncu -i X.ncu-rep --page raw --csv` wrote on the GPU box (the reports themselves
    # are too large to bring back: gpurun_out/ is limited to 64 MiB)
    if src.endswith(".csv"):
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")]
    with open(dst, "w") as f:
        f.write("# ncu --set full capture\n\nCommand: `%s`\n\nSource: `%s` (read with `ncu -i ... --page raw --csv`).\n\n" % (cmd, src))
        for r in rows[2:]:
            name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
            f.write("## `%s`\n\n| metric | value |\n|---|---|\n" % name)
            for m, label in FULL_METRICS:
                if m in col:
                    f.write("| %s (`%s`) | %s %s |\n" % (label, m, r[col[m]], units[col[m]]))
            st = []
            for h in stall_cols:
                try:
                    st.append((float(r[col[h]]), h))
                except ValueError:
                    pass
            st.sort(reverse=True)
            if st:
                f.write("\nTop stall reasons (warps stalled per issue): " +
                        ", ".join("%s %.2f" % (h.split("stalled_")[1].split("_per_issue")[0].replace(".ratio", ""), v) for v, h in st[:5]) + "\n")
            f.write("\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    {"launches": launches, "full": full}[mode](src, dst, cmd)
