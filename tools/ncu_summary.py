"""Turn ncu reports brought back in gpurun_out/ into the small CSV summaries committed under profiles/.
usage: python tools/ncu_summary.py out.csv label=path.ncu-rep[:launch_index] ...  [--metrics m1,m2,...]"""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
           "launch__grid_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    out, specs, metrics = sys.argv[1], [], DEFAULT
    for a in sys.argv[2:]:
        if a.startswith("--metrics"):
            metrics = a.split("=", 1)[1].split(",")
        else:
            specs.append(a)
    cols = []
    for sp in specs:
        label, path = sp.split("=", 1)
        idx = 0
        if ":" in path:
            path, i = path.rsplit(":", 1)
            idx = int(i)
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        h, units, row = rr[0], rr[1], rr[2 + idx]
        cols.append((label, row[h.index("Kernel Name")], {m: (units[h.index(m)], row[h.index(m)]) for m in metrics if m in h}))
    with open(out, "w") as f:
        for label, kname, _ in cols:
            f.write("# %s = %s\n" % (label, kname))
        f.write("metric,unit," + ",".join(c[0] for c in cols) + "\n")
        for m in metrics:
            unit = next((c[2][m][0] for c in cols if m in c[2]), "")
            # ncu picks a unit per report: a cell whose unit differs from the column header's carries its own
            cells = []
            for c in cols:
                u, v = c[2].get(m, ("", ""))
                v = v.replace(",", "")
                cells.append(v if (u == unit or not v) else v + " " + u)
            f.write(m + "," + unit + "," + ",".join(cells) + "\n")


if __name__ == "__main__":
    main()
