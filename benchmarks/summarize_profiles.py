#!/usr/bin/env python
"""Turn the raw gpurun_out/ artefacts of a round into the committed summaries under profiles/.

    python benchmarks/summarize_profiles.py r1

Reads (whatever exists):  gpurun_out/<tag>_launches_step.csv   ncu gpu__time_duration launch list of one bench step
                          gpurun_out/<tag>_instep_kernels.ncu-rep / <tag>_cold_kernels.ncu-rep   ncu --set full
                          gpurun_out/<tag>_sweep.jsonl          benchmarks/sweep_warp.py output
Writes: profiles/<tag>_bench_step_launches.md, profiles/<tag>_ncu_kernels.md, profiles/<tag>_sweep.md and
profiles/roofline_traffic.json (per-launch DRAM bytes of each hand-written kernel inside the bench step).
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank-conflict wavefronts"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]
OURS = ("warp_", "tlerp_", "act_tlerp_", "tok_", "mix_", "flow_head_", "bn_")      # kernel-name prefixes of libsmow_b200.so
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TO_US = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("smow::", "")


def launches(tag):
    p = os.path.join(OUT, tag + "_launches_step.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if not l.startswith("==")]
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        us = float(row["Metric Value"].replace(",", "")) * TO_US.get(row["Metric Unit"], 1)
        k = re.sub(r"<.*", "", short(row["Kernel Name"]))[:80]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    with open(os.path.join(PROF, tag + "_bench_step_launches.md"), "w") as f:
        f.write("# %s — kernels of ONE timed bench step (SMOW_Net_LW fwd+bwd, batch 16, 1x B200)\n\n" % tag)
        f.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off "
                "python bench.py --profile-step --steps 1 --warmup 3` (cold-cache, serialised: compare SHARES).\n\n")
        f.write("%d launches, %.1f ms of kernel time in total.\n\n| kernel | launches | us | share |\n|---|---:|---:|---:|\n" % (n, total / 1e3))
        ours = 0.0
        for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            mine = k.startswith(OURS)
            ours += us if mine else 0
            if us / total >= 0.003 or mine:
                f.write("| %s%s | %d | %.1f | %.2f %% |\n" % ("**" if mine else "", k + ("**" if mine else ""), c, us, 100 * us / total))
        f.write("\nHand-written kernels (rows A1-A5: warp_*, tlerp_*, act_tlerp_*; N1: flow_head_*; N2: tok_*; N4: mix_*, bn_*): %.1f us = "
                "%.2f %% of the step's kernel time.\n" % (ours, 100 * ours / total))


def ncu_raw(rep):
    """rows of an ncu report: either a .ncu-rep (converted here) or the raw-page CSV already written on the GPU box."""
    if rep.endswith(".csv"):
        text = open(rep).read()
    else:
        text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for row in rows[2:]:
        d = {"name": short(row[idx["Kernel Name"]])}
        for m, _ in METRICS:
            if m in idx:
                try:
                    v = float(row[idx[m]].replace(",", ""))
                except ValueError:
                    continue
                u = units[idx[m]]
                if u in TO_BYTES:
                    v *= TO_BYTES[u]
                elif m.startswith("gpu__time"):
                    v *= TO_US.get(u, 1)
                d[m] = v
        out.append(d)
    return out


def kernels(tag):
    traffic = {}
    with open(os.path.join(PROF, tag + "_ncu_kernels.md"), "w") as f:
        f.write("# %s — `ncu --set full --clock-control none` of the hand-written kernels\n\n" % tag)
        f.write("Raw reports stay in gpurun_out/ (scratch); this file is the committed summary. Durations under ncu are\n"
                "cold-cache and serialised; CUDA-event timings are in the sweep / bench files.\n")
        cold_title = ("HBM-cold: benchmarks/one_kernel.py --C 32 --H 128 --B 64 --layout ndhwc (fp32, 1.07 GB working set)" if tag == "r1" else
                      "HBM-cold: benchmarks/one_kernel_r2.py (C-ABI calls on fresh operands, B = 64, C = 32, 128 x 128 unless the row says "
                      "otherwise; second launch of each)")
        for part, title in (("cold", cold_title),
                            ("instep", "inside one bench step (SMOW_Net_LW, batch 16; operands partly L2-resident)")):
            rep = os.path.join(OUT, "%s_%s_kernels.ncu-rep" % (tag, part))
            if not os.path.exists(rep):
                rep = rep[:-len(".ncu-rep")] + ".csv"
            if not os.path.exists(rep) or os.path.getsize(rep) == 0:
                continue
            rows = ncu_raw(rep)
            f.write("\n## %s\n\n| # | kernel | %s |\n|---|---|%s\n" % (title, " | ".join(t for _, t in METRICS), "---:|" * len(METRICS)))
            rows = [d for d in rows if d["name"].startswith(OURS)]
            for i, d in enumerate(rows):
                cells = []
                for m, _ in METRICS:
                    v = d.get(m)
                    if v is None:
                        cells.append("-")
                    elif m.startswith("dram__bytes") or m.startswith("lts__t_bytes"):
                        cells.append("%.1f MB" % (v / 1e6))
                    elif m.startswith("gpu__time"):
                        cells.append("%.1f us" % v)
                    elif v >= 1e6:
                        cells.append("%.2f M" % (v / 1e6))
                    else:
                        cells.append("%.1f" % v if v != int(v) else "%d" % v)
                f.write("| %d | %s | %s |\n" % (i, d["name"][:60], " | ".join(cells)))
                if part == "instep":
                    key = re.sub(r"<.*", "", d["name"])
                    t = traffic.setdefault(key, [0, 0.0, 0.0])
                    b = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
                    t[0] += 1
                    t[1] += b
                    t[2] = max(t[2], b)
    if traffic:
        op_of = {"warp_fwd": "warp_stack_fwd", "warp_stack_fwd": "warp_stack_fwd", "warp_bwd": "warp_stack_bwd",
                 "warp_stack_bwd": "warp_stack_bwd", "tlerp_cat_fwd": "tlerp_cat_fwd", "tlerp_cat_bwd": "tlerp_cat_bwd",
                 "act_tlerp_cat_bwd": "tlerp_cat_bwd", "bn_act_bwd": "tlerp_cat_bwd", "bn_act_tlerp_cat_bwd": "tlerp_cat_bwd",
                 "tok_fwd": "tokenizer_fwd", "tok_bwd": "tokenizer_bwd",
                 "mix_apply": "frame_mix_apply", "mix_wgrad": "frame_mix_wgrad", "flow_head_fwd": "flow_head_fwd",
                 "flow_head_bwd": "flow_head_bwd"}
        per_op = {}
        for k, (n, b, bmax) in traffic.items():
            for pre, op in op_of.items():
                if k.startswith(pre):
                    o = per_op.setdefault(op, {"launches": 0, "bytes": 0.0, "largest": 0.0})
                    o["largest"] = max(o["largest"], bmax)
                    side = ("far", "combine", "bn_act_bwd_reduce", "bn_act_bwd_finalize", "bn_finalize", "pack", "w_reduce")
                    if not any(t in k for t in side):                                            # side kernels of the same C-ABI call
                        o["launches"] += n
                    o["bytes"] += b
                    break
        res = {op: v["bytes"] / max(1, v["launches"]) for op, v in per_op.items()}
        res.update({op + "@largest": v["largest"] for op, v in per_op.items()})      # the operator's biggest launch
        for plain, fused in (("tokenizer_fwd", "warp_tokens_fwd"), ("tokenizer_bwd", "warp_tokens_bwd")):
            if plain in res:              # in the bench step the tokenizer kernels run in their fused (row-staging) form
                res[fused], res[fused + "@largest"] = res[plain], res[plain + "@largest"]
        if "frame_mix_apply" in res:          # bench.py names the two uses of the apply kernel separately
            for alias in ("frame_mix_fwd", "frame_mix_bwd"):
                res[alias], res[alias + "@largest"] = res["frame_mix_apply"], res["frame_mix_apply@largest"]
        res["_source"] = "profiles/%s_ncu_kernels.md (in-step capture), dram__bytes_read+write per C-ABI call" % tag
        json.dump(res, open(os.path.join(PROF, "roofline_traffic.json"), "w"), indent=1)


def sweep(tag):
    p = os.path.join(OUT, tag + "_sweep.jsonl")
    if not os.path.exists(p):
        return
    rows = [json.loads(l) for l in open(p)]
    with open(os.path.join(PROF, tag + "_sweep.md"), "w") as f:
        f.write("# %s — isolated kernel sweep (BASELINE.json configs[4]), 1x B200, CUDA events, median of 10\n\n" % tag)
        f.write("`python benchmarks/sweep_warp.py --iters 10`; working set >= 1 GiB per launch (HBM-cold). GB/s = ALGORITHMIC bytes\n"
                "(SURVEY §8(d)) / time; frac = of the measured copy bandwidth (MEASURED_PEAKS.json, 6547.8 GB/s); ref = the reference's\n"
                "own op sequence (F.grid_sample + cat / F.interpolate + cat, ATen sm_100 kernels) on the same GPU.\n"
                "variant: 0 direct / atomics, 1 bulk-copy staged planes, 2 channel-vectorised tiles (NCDHW); 9 NDHWC (channels_last_3d)\n"
                "kernels as the modules run them (backward = tile gather + far pass), 8 NDHWC vector-atomic scatter backward.\n\n")
        f.write("| op | dtype | C | H=W | B | sigma | variant | ms | GB/s | frac | ref ms | speed-up |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows:
            if not r["op"].startswith(("warp", "tokenizer")):
                continue
            f.write("| %s | %s | %d | %d | %d | %.1f | %d | %.3f | %.0f | %.2f | %s | %s |\n" % (
                r["op"], r["dtype"], r["C"], r["H"], r["B"], r["sigma"], r["variant"], r["ms"], r["gbps"], r["frac"],
                "%.3f" % r["ref_ms"] if r["ref_ms"] else "-", "%.1fx" % r["speedup"] if r["speedup"] else "-"))
        f.write("\n| op | dtype | Cd | Cs | h=w | B | variant | ms | GB/s | frac | ref ms | speed-up |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|\n")
        for r in rows:
            if r["op"].startswith(("warp", "tokenizer")):
                continue
            f.write("| %s | %s | %d | %d | %d | %d | %d | %.3f | %.0f | %.2f | %.3f | %.1fx |\n" % (
                r["op"], r["dtype"], r["Cd"], r["Cs"], r["h"], r["B"], r["variant"], r["ms"], r["gbps"], r["frac"], r["ref_ms"], r["speedup"]))


def bench_blocks(tag):
    """profiles/<tag>_bench.md from gpurun_out/<tag>_bench.log (the builder's own full bench.py run): headline, roofline table,
    cfg3 / cfg4 / gpu_reference and the configs[4] sweep with its clock record."""
    p = os.path.join(OUT, tag + "_bench.log")
    if not os.path.exists(p):
        return
    line = [l for l in open(p) if l.startswith("{")]
    if not line:
        return
    d = json.loads(line[-1])
    with open(os.path.join(PROF, tag + "_bench.md"), "w") as f:
        f.write("# %s — builder-side `python bench.py` on 1x B200 (the driver's BENCH_rNN.json is the graded number)\n\n" % tag)
        f.write("* headline: **%.1f pairs/s** device-resident (%.2f ms/step), e2e %.1f pairs/s (%.2f ms/step, %d B H2D + %d B D2H per step); "
                "clocks %s\n" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"],
                                 d["e2e"]["d2h_bytes_per_step"], json.dumps(d.get("clocks"))))
        for k in ("cfg3", "cfg4", "gpu_reference", "cpu_baseline"):
            v = d.get(k)
            if isinstance(v, dict):
                f.write("* %s: `%s`\n" % (k, json.dumps({a: b for a, b in v.items() if a not in ("workload", "how")})[:900]))
        r = d.get("roofline")
        if isinstance(r, dict) and "all_kernels" in r:
            f.write("\n## roofline (graph-replayed C-ABI calls, HBM-cold rotation over >= 1 GiB; peak %.1f GB/s measured)\n\n" % r["peak"])
            f.write("dominant: `%s` %s: %.1f us, %.0f GB/s = **%.2f**; hot path = %.2f %% of the step\n\n" % (
                r["kernel"], json.dumps(r["shape"]), r["ms_per_launch"] * 1e3, r["achieved"], r["frac"], 100 * r["hot_path_share_of_step"]))
            f.write("| kernel | row | calls/step | ms/step cold | ms/step warm | algorithmic GB/s | frac |\n|---|---|---:|---:|---:|---:|---:|\n")
            for k, v in sorted(r["all_kernels"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
                f.write("| %s | %s | %.0f | %.3f | %.3f | %.0f | %.2f |\n" % (k, v["row"], v["calls_per_step"], v["ms_per_step"],
                                                                           v["ms_per_step_warm"], v["achieved_GB_per_s"], v["frac"]))
            f.write("\n| launch | shape | cold us | warm us | bytes | overhead bytes | frac |\n|---|---|---:|---:|---:|---:|---:|\n")
            for l in r["launches"]:
                f.write("| %s | `%s` | %.1f | %.1f | %d | %d | %.2f |\n" % (l["kernel"], json.dumps(l["shape"]), l["cold_ms"] * 1e3,
                                                                         l["warm_ms"] * 1e3, l["bytes_per_launch"], l["overhead_bytes"], l["frac"]))
        s_ = d.get("sweep")
        if isinstance(s_, dict) and "rows" in s_:
            f.write("\n## configs[4] sweep (clocks during the sweep: %s)\n\n%s\n\n" % (json.dumps(s_["clocks"]), s_["workload"]))
            f.write("summary (min / max fraction of the measured peak): `%s`\n\n" % json.dumps(s_["summary"]))
            f.write("| op | C | H=W | B | dtype | layout | sigma | ms | GB/s | frac | ATen fp32 NCDHW ms | speed-up |\n|---|---:|---:|---:|---|---|---:|---:|---:|---:|---:|---:|\n")
            for x in s_["rows"]:
                if "ms" in x:
                    f.write("| %s | %d | %d | %d | %s | %s | %.1f | %.3f | %.0f | %.2f | %.3f | %.1fx |\n" % (
                        x["op"], x["C"], x["H"], x["B"], x["dtype"], x["layout"], x["sigma"], x["ms"], x["GB_per_s"], x["frac"],
                        x["aten_f32_ncdhw_ms"], x["speedup_vs_aten"]))


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    kernels(tag)
    sweep(tag)
    bench_blocks(tag)
    print(sorted(os.listdir(PROF)))
