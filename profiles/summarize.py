#!/usr/bin/env python
"""Turn ncu output (gpurun_out/) into the small tracked summaries under profiles/.

    python profiles/summarize.py full  <X.ncu-rep | X_raw.csv>  <tag>  <workload>   # workload: scene0 | c4 | c5
        -> profiles/<tag>_<workload>_full.json/.md and the workload's entry of profiles/rollout_traffic.json
           (X_raw.csv / X_source.csv: `ncu -i X.ncu-rep --page raw|source --csv` exported on the GPU box when the
            report itself is too large to bring back)
    python profiles/summarize.py list  gpurun_out/launches_X.csv  <tag>  "<command that was profiled>"
        -> profiles/<tag>_launches.md and profiles/<tag>_launches_raw.csv

Runs here (no GPU): it only reads reports with `ncu -i` or the exported CSVs.
"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ENV_STEPS = (1 << 20) * 64
B_ALG = {"scene0": 22.0, "c4": 22.0, "c5": 22.5}
WORKLOAD = {"scene0": "scene_0 manual 9x9 map", "c4": "one 1024x1024 Bernoulli(0.002) map",
            "c5": "4096 distinct 256x256 Bernoulli(0.008) maps (one per 256 envs)"}

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ldgsts.sum",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
    "dram__sectors_read.sum", "dram__sectors_write.sum",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def read_page(path, page):
    if path.endswith(".ncu-rep"):
        out = subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True, check=True).stdout
        return list(csv.reader(io.StringIO(out)))
    p = path if page == "raw" else path.replace("_raw.csv", "_source.csv")
    if not os.path.exists(p):
        return None
    return list(csv.reader(open(p)))


def to_num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return v


def full(path, tag, workload):
    rows = read_page(path, "raw")
    head, units = rows[0], rows[1]
    d = dict(zip(head, rows[-1]))
    m = {"kernel": d.get("Kernel Name", "")}
    for k in RAW_KEYS:
        if k in d and d[k] != "":
            m[k] = {"value": to_num(d[k]), "unit": units[head.index(k)]}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    rd = m["dram__bytes_read.sum"]["value"] * scale.get(m["dram__bytes_read.sum"]["unit"], 1.0)
    wr = m["dram__bytes_write.sum"]["value"] * scale.get(m["dram__bytes_write.sum"]["unit"], 1.0)
    m["dram_bytes"] = rd + wr
    warp_steps = ENV_STEPS / 32.0
    total = m["smsp__inst_executed.sum"]["value"]
    summary = {"report": os.path.basename(path), "tag": tag, "workload": workload, "metrics": m,
               "warp_instructions_per_warp_step": round(total / warp_steps, 2)}
    src = read_page(path, "source")
    if src:
        hi = next(i for i, r in enumerate(src) if r and r[0] == "Address")
        h = src[hi]
        i_src, i_ex, i_sm, i_th = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
        inst = [(r[i_src].strip(), float(r[i_ex] or 0), float(r[i_sm] or 0), float(r[i_th] or 0)) for r in src[hi + 1:] if len(r) > i_th]
        tot_s = sum(x[2] for x in inst) or 1.0
        # 64-instruction regions with their share of instructions / stall samples / active threads
        regions = []
        for b0 in range(0, len(inst), 64):
            blk = inst[b0:b0 + 64]
            ins, smp, thr = sum(x[1] for x in blk), sum(x[2] for x in blk), sum(x[3] for x in blk)
            if ins / total > 0.01 or smp / tot_s > 0.01:
                regions.append({"sass_rows": [b0, b0 + len(blk)], "first": blk[0][0][:60], "inst_share": round(ins / total, 4),
                                "sample_share": round(smp / tot_s, 4), "threads_per_inst": round(thr / max(ins, 1), 1),
                                "inst_per_warp_step": round(ins / warp_steps, 2)})
        summary["sass_rows"] = len(inst)
        summary["regions"] = regions
    with open(os.path.join(HERE, "%s_%s_full.json" % (tag, workload)), "w") as f:
        json.dump(summary, f, indent=1)
    # keyed traffic file read by bench.py
    tpath = os.path.join(HERE, "rollout_traffic.json")
    try:
        traffic = json.load(open(tpath))
        if "dram_bytes_per_launch" in traffic:      # round-1 flat format
            traffic = {}
    except Exception:
        traffic = {}
    traffic[workload] = {"source": "profiles/%s_%s_full.json (ncu --set full, one launch, 2^20 envs x 64 steps, RECORD, %s)"
                                   % (tag, workload, WORKLOAD[workload]),
                         "dram_bytes_per_launch": m["dram_bytes"], "dram_bytes_read": rd, "dram_bytes_written": wr,
                         "algorithmic_bytes_per_launch": B_ALG[workload] * ENV_STEPS, "kernel": m["kernel"][:80]}
    with open(tpath, "w") as f:
        json.dump(traffic, f, indent=1)
    lines = ["# ncu --set full: %s (%s, %s)" % (m["kernel"].split("(")[0].replace("void ", ""), tag, os.path.basename(path)), "",
             "One launch = 2^20 envs x 64 env-steps, RECORD mode, FAST engine, %s.  Cold-cache, serialised" % WORKLOAD[workload],
             "replay: compare shares and counters, not absolute time.", "", "| metric | value | unit |", "|---|---|---|"]
    for k in RAW_KEYS:
        if k in m:
            lines.append("| `%s` | %s | %s |" % (k, m[k]["value"], m[k]["unit"]))
    lines += ["| dram bytes read+written per launch | %.4g | byte |" % m["dram_bytes"],
              "| algorithmic bytes per launch (%.1f B x 2^26 env-steps) | %.4g | byte |" % (B_ALG[workload], B_ALG[workload] * ENV_STEPS),
              "| traffic / algorithmic | %.3f | |" % (m["dram_bytes"] / (B_ALG[workload] * ENV_STEPS)),
              "| warp instructions per warp-step (32 env-steps) | %.1f | inst |" % (total / warp_steps), ""]
    if src:
        lines += ["| SASS rows | share of instructions | share of stall samples | threads / instruction | inst / warp-step | first instruction |",
                  "|---|---|---|---|---|---|"]
        for r in summary["regions"]:
            lines.append("| %d-%d | %.1f%% | %.1f%% | %.1f | %.1f | `%s` |" % (r["sass_rows"][0], r["sass_rows"][1], 100 * r["inst_share"],
                                                                    100 * r["sample_share"], r["threads_per_inst"], r["inst_per_warp_step"], r["first"]))
    with open(os.path.join(HERE, "%s_%s_full.md" % (tag, workload)), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


def launch_list(path, tag, command):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0] != "ID"]
    agg, order = {}, []
    for r in rows:
        name = r[4]
        short = name.split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        if "at::" in name:
            short = "torch:" + (name.split("at::")[1].split("<")[0])[:40]
        t = float(r[-1])
        if short not in agg:
            agg[short] = [0, 0.0, r[7], r[8]]
            order.append(short)
        agg[short][0] += 1
        agg[short][1] += t
    total = sum(v[1] for v in agg.values())
    lines = ["# ncu launch list (%s): gpu__time_duration.sum per kernel" % tag, "",
             "Command: `%s`." % command, "Cold-cache, serialised: shares only.", "",
             "| kernel | launches | block | grid | total us | share |", "|---|---|---|---|---|---|"]
    for k in sorted(order, key=lambda k: -agg[k][1]):
        n, t, blk, grd = agg[k]
        lines.append("| `%s` | %d | %s | %s | %.1f | %.1f%% |" % (k, n, blk, grd, t / 1e3, 100.0 * t / total))
    with open(os.path.join(HERE, "%s_launches.md" % tag), "w") as f:
        f.write("\n".join(lines) + "\n")
    shutil.copyfile(path, os.path.join(HERE, "%s_launches_raw.csv" % tag))
    print("\n".join(lines))


if __name__ == "__main__":
    mode, path, tag = sys.argv[1:4]
    if mode == "full":
        full(path, tag, sys.argv[4] if len(sys.argv) > 4 else "scene0")
    else:
        launch_list(path, tag, sys.argv[4] if len(sys.argv) > 4 else "python bench.py --steps 2 --warmup 3 --sub-steps 1 --no-e2e --no-cpu-baseline")
