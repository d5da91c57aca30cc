#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep / launches*.csv into the small tracked summaries under profiles/.

    python profiles/summarize.py full  gpurun_out/prof_rollout_X.ncu-rep  r1c   # -> profiles/<tag>_rollout_full.json/.md, rollout_traffic.json
    python profiles/summarize.py list  gpurun_out/launches_r1.csv         r1    # -> profiles/<tag>_launches.md

Runs here (no GPU): it only reads reports with `ncu -i`.
"""
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

RAW_KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
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
    "smsp__sass_average_branch_targets_threads_uniform.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
    "local_load_bytes", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def to_num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return v


def full(rep, tag, env_steps_per_launch=(1 << 20) * 64):
    rows = ncu_csv(rep, "raw")
    head, units = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        m = {"kernel": d.get("Kernel Name", "")}
        for k in RAW_KEYS:
            if k in d:
                m[k] = {"value": to_num(d[k]), "unit": units[head.index(k)]}
        launches.append(m)
    # per-launch DRAM traffic in bytes (ncu prints scaled units)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for m in launches:
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            if k in m:
                tot += m[k]["value"] * scale.get(m[k]["unit"], 1.0)
        m["dram_bytes"] = tot
    src = ncu_csv(rep, "source")
    h = src[1]
    i_src, i_ex, i_sm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    inst = [(r[i_src].strip(), int(r[i_ex]), int(r[i_sm])) for r in src[2:] if len(r) > i_ex]
    warp_steps = env_steps_per_launch / 32.0
    total = sum(x[1] for x in inst)
    hot = [{"i": i, "sass": s[:80], "per_warp_step": round(c / warp_steps, 3), "samples": sm}
           for i, (s, c, sm) in enumerate(inst) if c >= 0.2 * warp_steps]
    summary = {"report": os.path.basename(rep), "tag": tag, "launches": launches,
               "warp_instructions_per_warp_step": round(total / warp_steps, 2),
               "sass_rows": len(inst), "hot_loop_sass": hot}
    with open(os.path.join(HERE, "%s_rollout_full.json" % tag), "w") as f:
        json.dump(summary, f, indent=1)
    m = launches[-1]
    with open(os.path.join(HERE, "rollout_traffic.json"), "w") as f:
        json.dump({"source": "profiles/%s_rollout_full.json (ncu --set full, one launch of k_rollout, "
                             "2^20 envs x 64 steps, RECORD)" % tag,
                   "dram_bytes_per_launch": m["dram_bytes"], "kernel": m["kernel"][:60]}, f, indent=1)
    lines = ["# ncu --set full: k_rollout (%s, %s)" % (tag, os.path.basename(rep)), "",
             "One launch = 2^20 envs x 64 env-steps, RECORD mode, FAST engine, scene_0 grid.  Cold-cache, serialised",
             "replay: compare shares and counters, not absolute time.", "", "| metric | value | unit |", "|---|---|---|"]
    for k in RAW_KEYS:
        if k in m:
            lines.append("| `%s` | %s | %s |" % (k, m[k]["value"], m[k]["unit"]))
    lines += ["| dram bytes read+written per launch | %.4g | byte |" % m["dram_bytes"],
              "| algorithmic bytes per launch (22 B x 2^26 env-steps) | %.4g | byte |" % (22.0 * env_steps_per_launch),
              "| warp instructions per warp-step (32 env-steps) | %.1f | inst |" % (total / warp_steps), ""]
    with open(os.path.join(HERE, "%s_rollout_full.md" % tag), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


def launch_list(path, tag):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0] != "ID"]
    agg = {}
    order = []
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
             "Command: `python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline` (setup + 6 rollout launches).",
             "Cold-cache, serialised: shares only.", "", "| kernel | launches | block | grid | total us | share |",
             "|---|---|---|---|---|---|"]
    for k in sorted(order, key=lambda k: -agg[k][1]):
        n, t, blk, grd = agg[k]
        lines.append("| `%s` | %d | %s | %s | %.1f | %.1f%% |" % (k, n, blk, grd, t / 1e3, 100.0 * t / total))
    with open(os.path.join(HERE, "%s_launches.md" % tag), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    mode, path, tag = sys.argv[1:4]
    if mode == "full":
        full(path, tag)
    else:
        launch_list(path, tag)
