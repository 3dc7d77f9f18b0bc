#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<tag>_*.{csv,md,json}.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep TAG [--env-steps N] [--launches launches.csv]

Writes: <tag>_metrics.csv (key raw metrics), <tag>_hot_lines.md (instruction counts per source line / region and
SASS opcode mix of the top kernel), and updates profiles/traffic.json (dram bytes per launch, read by bench.py) when the
capture is of the headline step kernel, <tag>_traffic.json otherwise.
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


def kernel_source_hash():
    """sha256 over the CUDA sources of the library (same function as bench.py's): bench.py only claims these numbers
    for a library built from the very sources that were profiled.  Run this tool BEFORE editing the kernels again."""
    import hashlib

    h = hashlib.sha256()
    for f in ("qrmsa_kernels.cuh", "qrmsa_b200.cu"):
        h.update(open(os.path.join(ROOT, "optical_networking_gym_b200", "csrc", f), "rb").read())
    return h.hexdigest()


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
               "tag": tag, "kernel_source_sha256": kernel_source_hash()}
    if env_steps:
        traffic.update(env_steps_per_launch=env_steps, dram_bytes_per_env_step=dram / env_steps,
                       warp_instructions_per_env_step=inst / env_steps)

    # ---- exact totals from the SASS-only page (every instruction once): opcode mix, lane utilisation
    sass = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    sh = next(r for r in sass if len(r) > 5 and r[0] == "Address")
    iI, iT, iP = sh.index("Instructions Executed"), sh.index("Thread Instructions Executed"), sh.index("Predicated-On Thread Instructions Executed")
    iS = sh.index("# Samples")
    ops, per_addr, tot, thr, pon = collections.Counter(), {}, 0, 0, 0
    freq = collections.OrderedDict()   # runs of consecutive SASS instructions executed equally often = one loop level
    for r in sass:
        if len(r) <= iP or not r[0].startswith("0x"):
            continue
        n = int(r[iI]); tot += n; thr += int(r[iT]); pon += int(r[iP])
        p = r[1].split()
        op = (p[1] if p[0].startswith("@") else p[0]).split(".")[0]
        ops[op] += n
        per_addr[r[0]] = (n, int(r[iT]), int(r[iP]), int(r[iS] or 0))
        if env_steps and n:
            k = round(n / env_steps, 2)
            f = freq.setdefault(k, [0, 0])
            f[0] += 1; f[1] += n
    # ---- attribution to source lines from the CUDA+SASS page.  Inlined code is listed under its own line AND under
    # the lines of its callers, so every SASS address is counted ONCE, for the first line it appears under (the page
    # lists a file's lines in order, which puts a callee's own line before its call sites further down the file).
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]))))
    lines, lanes, text, seen = collections.Counter(), collections.Counter(), {}, set()
    samples = collections.Counter()
    fname, cur, infile = "", None, False
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            infile = r[1].endswith(".cuh") or r[1].endswith(".cu"); fname = os.path.basename(r[1]); cur = None; continue
        if len(r) < 8 or r[0] == "Line No":
            continue
        if r[0] != "":
            try:
                cur = (fname, int(r[0])); text[cur] = r[1].strip()
            except ValueError:
                cur = None
            continue
        if not r[2].startswith("0x") or r[2] in seen or r[2] not in per_addr:
            continue
        seen.add(r[2])
        if infile and cur:
            lines[cur] += per_addr[r[2]][0]
            lanes[cur] += per_addr[r[2]][1]
            samples[cur] += per_addr[r[2]][3]
    attributed = sum(lines.values())
    # ---- stages = the device function a line belongs to (definitions parsed from the source file)
    stage_of = {}
    try:
        import re
        path = os.path.join(ROOT, "optical_networking_gym_b200", "csrc", "qrmsa_kernels.cuh")
        name, struct = "(file scope)", None
        for i, l in enumerate(open(path), 1):
            ms = re.match(r"^(?:template.*>\s*)?struct\s+([A-Za-z_0-9]+)", l)
            if ms and l.rstrip().endswith("{"):
                struct = ms.group(1)
            if l.startswith("};"):
                struct = None
            mk = re.search(r"\b(k_[a-z_0-9]+)\s*\(", l)
            md = re.match(r"^(\s*)(?:template.*>\s*)?(?:static\s+)?(?:__host__\s+)?(?:__device__|__global__).*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(", l)
            if mk and ("__global__" in l or l.startswith("    k_")):
                name = mk.group(1)
            elif md and md.group(2) not in ("__launch_bounds__", "__align__"):
                name = (struct + "::" if struct and md.group(1) else "") + md.group(2)
            stage_of[("qrmsa_kernels.cuh", i)] = name
    except Exception:
        pass
    stages, stage_lanes, stage_samples = collections.Counter(), collections.Counter(), collections.Counter()
    for ln, n in lines.items():
        stages[stage_of.get(ln, ln[0])] += n
        stage_lanes[stage_of.get(ln, ln[0])] += lanes[ln]
        stage_samples[stage_of.get(ln, ln[0])] += samples[ln]
    all_samples = max(sum(v[3] for v in per_addr.values()), 1)
    per = env_steps or 1.0
    traffic.update(warp_instructions_per_launch_sass_page=tot, avg_threads_per_instruction=thr / max(tot, 1),
                   avg_predicated_on_threads_per_instruction=pon / max(tot, 1))
    # bench.py reads profiles/traffic.json for the HEADLINE step kernel only: captures of other kernels keep their own file
    is_headline = "k_step_policy<320" in traffic["kernel"] or "k_step_policy<(int)320" in traffic["kernel"]
    json.dump(traffic, open(os.path.join(out_dir, "traffic.json" if is_headline else f"{tag}_traffic.json"), "w"), indent=1)
    with open(os.path.join(out_dir, f"{tag}_hot_lines.md"), "w") as f:
        f.write(f"# {tag}: {m.get('Kernel Name', ('', ''))[0]}\n\n")
        f.write(f"warp-instructions per launch {tot:.4g}" + (f" = {tot / per:.1f} per env-step" if env_steps else "") +
                f" (SASS page, every instruction once; smsp__inst_executed.sum = {inst:.4g}); "
                f"{thr / max(tot, 1):.2f} threads active per instruction, {pon / max(tot, 1):.2f} predicated on; "
                f"{100.0 * attributed / max(tot, 1):.1f} % attributed to source lines below\n\n")
        f.write("## SASS opcode mix (per env-step)\n\n" if env_steps else "## SASS opcode mix\n\n")
        f.write(", ".join(f"{o} {n / per:.1f}" for o, n in ops.most_common(30)) + "\n\n")
        if env_steps:
            f.write("## by execution frequency (SASS instructions that run equally often per env-step belong to one loop level)\n\n"
                    "| executions per env-step | SASS instructions | warp-instructions per env-step |\n|---|---|---|\n")
            for k, (cnt, n) in sorted(freq.items(), key=lambda kv: -kv[1][1])[:28]:
                f.write(f"| {k:.2f} | {cnt} | {n / per:.1f} |\n")
            f.write("\n")
        f.write("## stages (device function owning the line; warp-instructions per env-step, active lanes per instruction)\n\n"
                "| function | instr | lanes | % of stall samples |\n|---|---|---|---|\n")
        for st, n in stages.most_common(40):
            f.write(f"| {st} | {n / per:.1f} | {stage_lanes[st] / max(n, 1):.1f} | {100.0 * stage_samples[st] / all_samples:.1f} |\n")
        f.write("\n## source lines by warp-stall samples (where the time goes)\n\n| line | % samples | instr | source |\n|---|---|---|---|\n")
        for ln, n in samples.most_common(40):
            f.write(f"| {ln[0].replace('qrmsa_', '').replace('.cuh', '')}:{ln[1]} | {100.0 * n / all_samples:.1f} | {lines[ln] / per:.1f} | `{text[ln][:110]}` |\n")
        f.write("\n## hottest source lines (warp-instructions per env-step, active lanes per instruction)\n\n| line | instr | lanes | source |\n|---|---|---|---|\n")
        for ln, n in lines.most_common(70):
            f.write(f"| {ln[0].replace('qrmsa_', '').replace('.cuh', '')}:{ln[1]} | {n / per:.1f} | {lanes[ln] / max(n, 1):.1f} | `{text[ln][:110]}` |\n")
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
