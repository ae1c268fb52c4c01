"""profiles/summarize.py -- turn an ncu report into the small tracked summaries under profiles/.

  python profiles/summarize.py gpurun_out/prof_r01.ncu-rep r01 [gpurun_out/launches_r01.csv]

Writes profiles/ncu_<tag>_kernels.csv (one row per profiled launch, selected metrics from
`ncu --page raw --csv`), profiles/launches_<tag>_summary.csv (per-kernel totals of the
gpu__time_duration launch list) and, when the covproj kernel is present,
profiles/covproj_traffic.json (DRAM bytes per eval, read by bench.py for roofline.traffic).
"""
import collections
import csv
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum",
]


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    kn = hdr.index("Kernel Name")
    out = os.path.join(HERE, f"ncu_{tag}_kernels.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + [f"{m} [{units[i]}]" for m, i in cols])
        for r in data:
            name = r[kn].split("(")[0].replace("void ", "").replace("xyzb::<unnamed>::", "").replace("unnamed>::", "")
            w.writerow([name] + [r[i] for _, i in cols])
    print("wrote", out)
    rd, wr, t = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    cov = [r for r in data if "covproj_tma" in r[kn]]
    if cov:
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        per = [(float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]) for r in cov]
        elems = 1 << 26
        with open(os.path.join(HERE, "covproj_traffic.json"), "w") as f:
            json.dump({"dram_bytes_per_launch": sum(per) / len(per), "elems_per_launch": elems,
                       "dram_bytes_per_eval": sum(per) / len(per) / elems,
                       "source": f"ncu --set full, profiles/ncu_{tag}_kernels.csv (dram__bytes_read.sum + dram__bytes_write.sum, "
                                 f"{len(per)} launches of 2^26 elements)"}, f, indent=1)
    if len(sys.argv) > 3:
        rows = list(csv.reader(open(sys.argv[3])))
        h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
        hdr = rows[h]
        kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        to_ns = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}
        agg = collections.OrderedDict()
        for r in rows[h + 1:]:
            if len(r) > mv:
                agg.setdefault(r[kn].split("(")[0][-70:], []).append(float(r[mv].replace(",", "")) * to_ns.get(r[mu], 1.0))
        tot = sum(sum(v) for v in agg.values())
        out = os.path.join(HERE, f"launches_{tag}_summary.csv")
        with open(out, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["kernel", "launches", "total_us", "avg_us", "share_pct"])
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
                w.writerow([k, len(v), round(sum(v) / 1e3, 2), round(sum(v) / len(v) / 1e3, 3), round(100 * sum(v) / tot, 2)])
        print("wrote", out)


if __name__ == "__main__":
    main()
