"""profiles/r2_sort_ncu.txt from the captures of tools/ncu_final.sh (gpurun_out/r2_digits_pre.ncu-rep, r2_scatter_pre.ncu-rep):
what bounds the two passes of the counting sort on the precomputed path.   usage: python tools/ncu_sort_extract.py"""
import csv, io, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.max.pct_of_peak_sustained_elapsed",
        "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed", "lts__xbar2lts_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_sectors_mem_global_op_atom.sum", "l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return {n: (v, u) for n, u, v in zip(rows[0], rows[1], rows[2])}


def main():
    lines = ["Sort stage of the precomputed path at 2^20 terms, c = 18 (15.7 M entries, 131072 buckets): ncu --set full --clock-control none,",
             "one launch each, tools/ncu_final.sh (python tools/msm_probe.py --lgn 20 --pre 0), exact counting sort (slot sort off).", ""]
    for rep, title in (("r2_digits_pre.ncu-rep", "k_digits_pre  -- histogram pass: 32 B scalar read + 15 scattered RED.ADD per term"),
                       ("r2_scatter_pre.ncu-rep", "k_scatter_pre -- scatter pass: 15 returning ATOM.ADD + 15 scattered 8-byte stores per term")):
        p = os.path.join(ROOT, "gpurun_out", rep)
        if not os.path.exists(p):
            lines.append("missing " + rep); continue
        d = raw(p)
        lines.append(title)
        for k in KEYS:
            if k in d:
                lines.append("  %-88s %s %s" % (k, d[k][0], d[k][1]))
        lines.append("")
    lines += ["Reading: neither pass is bound by bytes (DRAM 4-14 %, coalesced traffic is small) nor by an L2 slice limit (lts 61-63 % on average,",
              "the busiest slice 74-88 %); both run at the same rate of SCATTERED requests per SM: 15.7 M RED in 111 us = 0.48 requests/clk/SM,",
              "15.7 M ATOM + 15.7 M stores in 232 us = 0.47 requests/clk/SM (148 SMs, 1.965 GHz).  The scatter pass waits on the returned",
              "slot numbers (long scoreboard 137 per issue), the histogram pass on the LSU queue (lg_throttle 26).  Time = scattered requests",
              "per entry x 7.4 ps: three for the exact counting sort; the slot sort (msm.cuh, k_scatter_slots_pre) needs two."]
    open(os.path.join(ROOT, "profiles", "r2_sort_ncu.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
