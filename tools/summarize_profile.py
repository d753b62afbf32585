#!/usr/bin/env python
"""Turn an ncu launch list (+ optional --set full report) into the markdown summary kept under profiles/.

    python tools/summarize_profile.py <tag> [--rep gpurun_out/<tag>_prof.ncu-rep]
reads gpurun_out/<tag>_launches.csv and gpurun_out/<tag>_bench.json, writes profiles/<tag>_summary.md
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RAW_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.max.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
            "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "lts__t_sector_hit_rate.pct",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(path):
    lines = open(path).read().splitlines()
    i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
    return list(csv.DictReader(lines[i:]))


def main():
    tag = sys.argv[1]
    reps = []
    if "--rep" in sys.argv:
        reps = sys.argv[sys.argv.index("--rep") + 1:]
    out = ["# ncu summary `%s`" % tag, ""]
    bj = os.path.join(ROOT, "gpurun_out", tag + "_bench.json")
    if os.path.exists(bj) and os.path.getsize(bj):
        b = json.loads(open(bj).read().strip().splitlines()[-1])
        out += ["Bench line of the same build (not under ncu): value %.4g %s, e2e %.4g, %.3f ms/step, launches %s" %
                (b["value"], b["unit"], b["e2e"]["value"], b["ms_per_step"], b.get("gpu_launches")), ""]
        out += ["stage ms per batch (CUDA events, concurrent batches): " +
                ", ".join("%s %.3f" % kv for kv in (b["roofline"].get("stage_ms_per_batch_concurrent") or b["roofline"].get("stage_ms_per_batch", {})).items()), ""]
    lp = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
    if os.path.exists(lp):
        rows = launches(lp)
        agg = collections.OrderedDict()
        for r in rows:
            k = r["Kernel Name"].split("(")[0].replace("void ", "")
            a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
            a[0] += 1
            a[1] += float(r["Metric Value"]) / 1e6
        tot = sum(v[1] for v in agg.values())
        out += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, `bench.py --steps 1 --warmup 1`)",
                "", "Serialised, cold-cache per-launch times: compare shares, not absolutes.", "",
                "| kernel | launches | total ms | avg ms | share | grid | block |", "|---|---|---|---|---|---|---|"]
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append("| `%s` | %d | %.3f | %.4f | %.1f%% | %s | %s |" % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot, v[2], v[3]))
        out.append("")
    for rep in reps:
        if not os.path.exists(rep):
            continue
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        out += ["## `ncu --set full` capture (%s)" % os.path.basename(rep), ""]
        kcol = hdr.index("Kernel Name")
        out += ["kernels: " + ", ".join("`%s`" % r[kcol].split("(")[0] for r in rows[2:]), "",
                "| metric | unit | " + " | ".join("launch %d" % i for i in range(len(rows) - 2)) + " |",
                "|---|---|" + "---|" * (len(rows) - 2)]
        for key in RAW_KEYS:
            if key in hdr:
                i = hdr.index(key)
                out.append("| %s | %s | %s |" % (key, units[i], " | ".join(r[i] for r in rows[2:])))
        out.append("")
    # DRAM traffic per launch of every captured kernel (bench.py's roofline.traffic reads the newest *_traffic.json)
    traffic = {}
    for rep in reps:
        if not os.path.exists(rep):
            continue
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")
            ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
            traffic[name] = {"dram_read_bytes": float(r[ir]) * scale.get(units[ir], 1.0), "dram_write_bytes": float(r[iw]) * scale.get(units[iw], 1.0),
                             "duration_us": float(r[it]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[it], 1.0),
                             "source": os.path.basename(rep) + " (ncu --set full, one launch of a 256 x 4000-sample batch)"}
    if traffic:
        json.dump(traffic, open(os.path.join(ROOT, "profiles", tag + "_traffic.json"), "w"), indent=1)
    dst = os.path.join(ROOT, "profiles", tag + "_summary.md")
    open(dst, "w").write("\n".join(out) + "\n")
    print(open(dst).read())


if __name__ == "__main__":
    main()
