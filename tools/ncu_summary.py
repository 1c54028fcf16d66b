"""Condense ncu output into the small text summaries committed under profiles/.

  python tools/ncu_summary.py rep   <file.ncu-rep> <out.txt>      key counters of every kernel in a --set full report
  python tools/ncu_summary.py list  <launches.csv> <out.txt>      per-kernel device time and share of one step
                                                                  (or of the whole list when it holds no hot-path step)
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none summary of %s\n" % path.split("/")[-1])
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("\n## %s  grid=%s block=%s\n" % (d.get("Kernel Name", "?")[:100], d.get("Grid Size"), d.get("Block Size")))
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write("%-95s %-14s %s\n" % (k, units[hdr.index(k)], d[k]))


def launch_list(path, out):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hi + 1:] if len(r) > vi and r[vi] not in ("", "Metric Value")
           and r[0] != ""]
    # one steady-state step = from the last-but-one framing kernel to the last one; any other launch list (e.g.
    # tools/attn_bench.py) is summarised whole
    idx = [i for i, (n, _) in enumerate(seq) if "pad_split" in n or "fold_split" in n]
    step = seq[idx[-2]:idx[-1]] if len(idx) >= 2 else seq
    tot = sum(v for _, v in step)
    agg = collections.OrderedDict()
    for n, v in step:
        short = n.split("(")[0][-70:]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: kernels of ONE steady-state step\n")
        f.write("# (cold-cache, serialised: compare shares, not absolutes).  step total = %.1f us, %d launches\n" % (tot / 1e3, len(step)))
        f.write("%-72s %6s %10s %7s\n" % ("kernel", "calls", "us", "share"))
        for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s %6d %10.1f %6.1f%%\n" % (n, c, v / 1e3, 100 * v / tot))


if __name__ == "__main__":
    {"rep": rep, "list": launch_list}[sys.argv[1]](sys.argv[2], sys.argv[3])
