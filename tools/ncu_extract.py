"""Turns the captures of tools/ncu_metrics.sh (gpurun_out/r2_*.ncu-rep, r2_launches_bench.csv) into the committed evidence:
profiles/r2_ncu_metrics.json (read by bench.py for roofline.traffic / fmaheavy %), profiles/r2_accumulate_ncu.txt,
profiles/r2_rp_lookup16_ncu.txt and profiles/r2_launches_bench.csv.   usage: python tools/ncu_extract.py"""
import csv, io, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "s": 1e9, "ms": 1e6, "us": 1e3, "ns": 1.0}      # bytes; times in ns


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for n, u, v in zip(hdr, units, vals):
        try:
            x = float(v.replace(",", ""))
        except ValueError:
            continue
        d[n] = x * UNIT[u] if u in UNIT else x
    d["_kernel"] = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    return d


def summary(rep, title, path):
    d = raw(rep)
    with open(path, "w") as f:
        f.write("%s\nkernel: %s\nsource: ncu --set full --clock-control none --import-source on (tools/ncu_metrics.sh), one launch, cold L2\n\n" % (title, d["_kernel"]))
        for k in KEYS:
            if k in d:
                f.write("%-92s %s\n" % (k, ("%.0f" % d[k]) if d[k] > 1e6 else ("%.3f" % d[k])))
    return d


def main():
    os.makedirs(PR, exist_ok=True)
    rec = {}
    try:      # captures that are no longer in gpurun_out/ keep their committed record
        rec = json.load(open(os.path.join(PR, "r2_ncu_metrics.json")))
    except Exception:   # noqa: BLE001
        pass
    rec["source"] = ("profiles/r2_accumulate_slots_ncu.txt, profiles/r2_accumulate_ncu.txt, profiles/r2_rp_lookup16_ncu.txt "
                     "(tools/ncu_metrics.sh, tools/ncu_final.sh + tools/ncu_extract.py)")
    for name, rep, title, txt in (("k_accumulate_slots", "r2_accumulate_slots.ncu-rep", "k_accumulate_slots, precomputed-window path after the slot sort, 2^20 terms, c = 18 (the bench's dominant kernel)", "r2_accumulate_slots_ncu.txt"),
                                  ("k_accumulate", "r2_accumulate.ncu-rep", "k_accumulate, precomputed-window path, 2^20 terms, c = 18 (compact sorted list; the dominant kernel before the slot sort)", "r2_accumulate_ncu.txt"),
                                  ("k_rp_lookup16", "r2_lookup16.ncu-rep", "k_rp_lookup16, 16-bit generator table, one chunk of an 8192-proof batch", "r2_rp_lookup16_ncu.txt")):
        p = os.path.join(GO, rep)
        if not os.path.exists(p):
            print("missing", p); continue
        d = summary(p, title, os.path.join(PR, txt))
        rec[name] = {"dram_bytes": int(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)),
                     "dram_bytes_read": int(d.get("dram__bytes_read.sum", 0)), "dram_bytes_write": int(d.get("dram__bytes_write.sum", 0)),
                     "pipe_fmaheavy_active_pct": d.get("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                     "issue_active_pct": d.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "warps_active_pct": d.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                     "registers": d.get("launch__registers_per_thread"), "duration_us": d.get("gpu__time_duration.sum", 0) / 1e3,
                     "l2_hit_pct": d.get("lts__t_sector_hit_rate.pct")}
    src = os.path.join(GO, "r2_launches_bench.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(PR, "r2_launches_bench.csv"))
        rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
        tot = {}
        for r in rows:
            k = r[4].split("(")[0]
            tot.setdefault(k, [0, 0.0]); tot[k][0] += 1; tot[k][1] += float(r[-1].replace(",", ""))
        rec["launch_list"] = {"file": "profiles/r2_launches_bench.csv", "command": "python bench.py --steps 2 --warmup 1 --no-verify --strong ''",
                              "kernels": {k: {"launches": v[0], "total_us": round(v[1] / 1e3, 1)} for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])}}
    with open(os.path.join(PR, "r2_ncu_metrics.json"), "w") as f:
        json.dump(rec, f, indent=1)
    print(json.dumps(rec, indent=1)[:3000])


if __name__ == "__main__":
    main()
