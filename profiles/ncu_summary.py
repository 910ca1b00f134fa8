"""Summarise an .ncu-rep (read here, no GPU): python profiles/ncu_summary.py <file.ncu-rep> [kernel-regex]"""
import collections
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for r in data:
        print("== kernel:", r[hdr.index("Kernel Name")][:90])
        for w in WANT:
            if w in hdr:
                print("  %-62s %14s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))


def source(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, sect = None, []
    for r in rows:
        if len(r) > 5 and r[0] == "Address":
            if hdr is not None:
                break
            hdr = r
            continue
        if hdr is not None and len(r) == len(hdr):
            sect.append(r)
    if hdr is None:
        return
    ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    tot = sum(int(r[ia]) for r in sect)
    print("  static SASS instr %d, executed warp-instr %d" % (len(sect), tot))
    buckets = collections.Counter()
    for r in sect:
        buckets[int(r[ia])] += 1
    for k, v in sorted(buckets.items(), reverse=True)[:8]:
        print("    executed %10d times: %4d instructions" % (k, v))
    stalls = {h: sum(int(r[i] or 0) for r in sect) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
    s = sum(stalls.values()) or 1
    print("  stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / s) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:7]))
    for r in sorted(sect, key=lambda r: -int(r[isamp]))[:6]:
        print("    hot: %6s samples  x%-9s %s" % (r[isamp], r[ia], r[isrc].strip()[:80]))


if __name__ == "__main__":
    raw(sys.argv[1])
    source(sys.argv[1])
