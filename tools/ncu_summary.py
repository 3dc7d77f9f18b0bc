#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<tag>_*.{csv,md,json}.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep TAG [--env-steps N] [--launches launches.csv]

Writes: <tag>_metrics.csv (key raw metrics), <tag>_hot_lines.md (instruction counts per source line / region and
SASS opcode mix of the top kernel), and updates profiles/traffic.json (dram bytes per launch, read by bench.py).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.per_cycle_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    env_steps = None
    if "--env-steps" in sys.argv:
        env_steps = float(sys.argv[sys.argv.index("--env-steps") + 1])
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    m = {}
    with open(os.path.join(out_dir, f"{tag}_metrics.csv"), "w") as f:
        f.write("metric,unit,value\n")
        for i, h in enumerate(hdr):
            if h in KEEP or h == "Kernel Name":
                f.write(f"{h},{units[i]},{vals[i]}\n")
                m[h] = (vals[i], units[i])

    def num(k):
        v, u = m[k]
        v = float(v.replace(",", ""))
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
        return v * scale.get(u, 1.0)

    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    dur = num("gpu__time_duration.sum")
    inst = num("smsp__inst_executed.sum")
    traffic = {"kernel": m.get("Kernel Name", ("", ""))[0], "dram_bytes_per_launch": dram,
               "duration_s_under_ncu": dur, "warp_instructions_per_launch": inst, "source": os.path.basename(rep),
               "tag": tag}
    if env_steps:
        traffic.update(env_steps_per_launch=env_steps, dram_bytes_per_env_step=dram / env_steps,
                       warp_instructions_per_env_step=inst / env_steps)
    json.dump(traffic, open(os.path.join(out_dir, "traffic.json"), "w"), indent=1)

    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    ops, lines, text = collections.Counter(), collections.Counter(), {}
    fname = ''
    stall = collections.Counter()
    infile, cur, tot, hdr_row = False, None, 0, None
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            infile = r[1].endswith(".cuh"); fname = os.path.basename(r[1]); cur = None; continue
        if len(r) >= 8 and r[0] == "Line No":
            hdr_row = r; continue
        if len(r) < 8:
            continue
        if r[0] != "":
            try:
                cur = (fname, int(r[0])); text[cur] = r[1].strip()
            except ValueError:
                cur = None
            continue
        try:
            n = int(r[7])
        except ValueError:
            continue
        p = r[3].split()
        if not p:
            continue
        op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
        ops[op] += n; tot += n
        if infile and cur:
            lines[cur] += n
    per = env_steps or 1.0
    with open(os.path.join(out_dir, f"{tag}_hot_lines.md"), "w") as f:
        f.write(f"# {tag}: {m.get('Kernel Name', ('', ''))[0]}\n\n")
        f.write(f"warp-instructions per launch {tot:.4g}" + (f" = {tot / per:.0f} per env-step" if env_steps else "") + "\n\n")
        f.write("## SASS opcode mix (per env-step)\n\n" if env_steps else "## SASS opcode mix\n\n")
        f.write(", ".join(f"{o} {n / per:.1f}" for o, n in ops.most_common(30)) + "\n\n")
        f.write("## hottest source lines (warp-instructions per env-step, csrc/*.cuh)\n\n| line | instr | source |\n|---|---|---|\n")
        for ln, n in lines.most_common(70):
            f.write(f"| {ln[0].replace('qrmsa_', '').replace('.cuh', '')}:{ln[1]} | {n / per:.1f} | `{text[ln][:110]}` |\n")
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
