#!/usr/bin/env python
"""bench.py — image-pairs/s of SMOW_Net_LW fwd+bwd (256x256, batch 16 per GPU) + hot-path roofline, plus the other
BASELINE.json configurations as extra blocks of the same JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Contract: W untimed warm-up steps, exactly K timed steps bracketed by
barrier + synchronize, CUDA-event timing, MAX over ranks, one JSON line from rank 0.
  value          whole-job pairs/s with inputs resident in HBM (BASELINE.json configs[1], weak-scaled at N > 1)
  e2e            same metric through the public module call with pinned HOST inputs (H2D inside the timed region,
                 double-buffered on a copy stream) and a D2H read of the loss every step
  roofline       the DOMINANT hand-written kernel of the step (largest time per step over ALL of them, rows A1-A5 and
                 N2 / N4 alike): SURVEY §8(d) algorithmic bytes / graph-replayed HBM-cold launch time vs the measured peak
  cpu_baseline   the reference's own modules (oracle/_ref/reference.zip) on the host cores (bounded sample)
  cfg3           BASELINE.json configs[2]: SMOW_Net FULL training step (fwd, BCE-Dice, bwd, clamp-clip, AdamW), global
                 batch 128 => 128/N pairs per GPU, one CUDA graph per step
  cfg4           BASELINE.json configs[3]: 1024x1024 tiles as 16 crops of 256x256, eval, sharded over the ranks, no collectives
  sweep          BASELINE.json configs[4] (N = 1 only): warp+stack fwd / bwd at C x H = {64,128,256} x {64,128,256}, fp32 + bf16,
                 sigma 0.3 / 8, NDHWC + NCDHW, HBM-cold, graph-replayed, with ATen's grid_sample+cat on the same GPU
  gpu_reference  (N = 1 only) the reference's own module on the same GPU the way train.py runs it (eager, NCDHW)
--impl reference times the reference's CPU implementation alone.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image-pairs/s 256x256 fwd+bwd"
UNIT = "pairs/s"
MODEL = "lw"
BATCH = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="pairs per GPU")
    ap.add_argument("--model", default=MODEL, choices=["lw", "s"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--strict-fp32", action="store_true", help="disable cuDNN TF32 convolutions (and the TF32 tensor-core frame mix)")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--skip", default="", help="comma list of extra blocks to skip: cfg3,cfg4,sweep,gpu_reference,roofline")
    ap.add_argument("--cfg3-global-batch", type=int, default=128)
    ap.add_argument("--profile-step", action="store_true",
                    help="ncu helper: warm up, then run --steps steps between cudaProfilerStart/Stop and exit")
    return ap.parse_args()


def config(args, world, launch_mode="eager"):
    name = "SMOW_Net_LW" if args.model == "lw" else "SMOW_Net"
    return {"launch": launch_mode, "workload": "%s forward+backward (BCE-Dice loss), batch %d per GPU, 256x256 synthetic pairs "
                        "(BASELINE.json configs[1])" % (name, args.batch),
            "global_batch": args.batch * world, "image_size": 256,
            "parallelism": "dp%d (DDP, NCCL gradient all-reduce)" % world if world > 1 else "single GPU",
            "l2": "per-step working set (activations+gradients, >1 GB) exceeds the 126 MB L2; no explicit flush",
            "conv_math": "fp32 hot path; cuDNN convolutions and the tcgen05 frame mix under torch defaults "
                         "(cudnn.allow_tf32=%s)" % (not args.strict_fp32)}


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampling during the timed region (profiling recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def ncu_traffic(kernel):
    """per-launch DRAM bytes of `kernel` from the committed ncu --set full summary, if one exists."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get(kernel, d.get(kernel + "@largest"))
    return None


def host_info():
    model = "unknown"
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                model = line.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return {"nproc": os.cpu_count(), "cpu_model": model}


def cpu_reference(kind, batch, steps, warmup, budget_s):
    """The reference's CPU implementation: its own modules from oracle/_ref/reference.zip when the archive travelled,
    else the oracle's port (oracle/cpu_model.py)."""
    from oracle import build_ref, ref_runtime
    build_ref.build()
    if ref_runtime.available():
        r = ref_runtime.time_cpu_fwd_bwd(kind, batch, steps, warmup, budget_s=budget_s)
        return r, "reference", "the reference's own %s (oracle/_ref/reference.zip, unmodified files)" % (
            "SMOW_Net_LW" if kind == "lw" else "SMOW_Net")
    from oracle import cpu_model
    r = cpu_model.time_cpu_fwd_bwd(kind, batch, steps, warmup, budget_s=budget_s)
    return r, "port", "the oracle's port of the reference model (reference archive absent)"


def run_reference(args, rank, world):
    if rank != 0:
        return
    r, kind, what = cpu_reference(args.model, args.batch, args.steps, args.warmup, 150.0)
    cb = dict({"value": r["pairs_per_s"], "unit": UNIT, "cores": r["threads"], "kind": kind,
               "sample": "%d pairs per step x %d steps, fwd+loss+bwd of %s on the host CPU" % (r["batch"], r["steps"], what)},
              **host_info())
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": r["pairs_per_s"], "unit": UNIT,
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": config(args, 1), "cpu_baseline": cb,
                      "e2e": {"value": r["pairs_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------- extra blocks
def guarded(fn):
    """Extra blocks never take the headline down: an exception becomes {"error": ...}."""
    try:
        return fn()
    except Exception as e:      # noqa: BLE001
        import traceback
        traceback.print_exc(file=sys.stderr)
        return {"error": "%s: %s" % (type(e).__name__, str(e).splitlines()[0][:200] if str(e) else "")}


def roofline_block(calls, step_ms, device, n_steps_recorded):
    """Every hand-written kernel of one step, rebuilt from its recorded shape and timed graph-replayed (cold = operands
    rotating over >= 1 GiB; warm = one operand set).  Dominant = the largest (calls per step x cold time) over ALL rows."""
    from smow_net_b200 import probe
    peak, peak_src = measured_peak()
    groups = {}
    for name, meta in calls:
        if meta is None:
            continue
        key = (name, tuple(sorted(meta.items())))
        g = groups.setdefault(key, {"name": name, "meta": meta, "calls": 0})
        g["calls"] += 1
    rows = []
    for g in groups.values():
        t = probe.time_call(g["name"], g["meta"], device, footprint=1 << 30, max_sets=48)
        per_step = g["calls"] / n_steps_recorded
        rows.append({"kernel": g["name"], "shape": {k: v for k, v in g["meta"].items()}, "calls_per_step": per_step,
                     "cold_ms": t["cold_ms"], "warm_ms": t["warm_ms"], "bytes_per_launch": t["bytes"],
                     "overhead_bytes": t["operand_bytes"] - t["bytes"],
                     "achieved_GB_per_s": t["bytes"] / t["cold_ms"] / 1e6, "frac": t["bytes"] / t["cold_ms"] / 1e6 / peak,
                     "ms_per_step_cold": per_step * t["cold_ms"], "ms_per_step_warm": per_step * t["warm_ms"]})
    per_kernel = {}
    for r in rows:
        k = per_kernel.setdefault(r["kernel"], {"ms_per_step": 0.0, "bytes_per_step": 0.0, "calls_per_step": 0.0, "ms_per_step_warm": 0.0})
        k["ms_per_step"] += r["ms_per_step_cold"]
        k["ms_per_step_warm"] += r["ms_per_step_warm"]
        k["bytes_per_step"] += r["bytes_per_launch"] * r["calls_per_step"]
        k["calls_per_step"] += r["calls_per_step"]
    for name, k in per_kernel.items():
        k["achieved_GB_per_s"] = k["bytes_per_step"] / k["ms_per_step"] / 1e6
        k["frac"] = k["achieved_GB_per_s"] / peak
        k["row"] = ("A1+N2" if name.startswith("warp_tokens") else "N2" if name.startswith("tokenizer") else "N4" if name.startswith("frame_mix") else
                    "N1" if name.startswith("flow_head") else "A1-A5")
    dom_name = max(per_kernel, key=lambda n: per_kernel[n]["ms_per_step"])
    dom_rows = [r for r in rows if r["kernel"] == dom_name]
    dom = max(dom_rows, key=lambda r: r["ms_per_step_cold"])
    total = sum(k["ms_per_step"] for k in per_kernel.values())
    return {"bound": "hbm", "kernel": dom_name, "achieved": dom["achieved_GB_per_s"], "peak": peak, "unit": "GB/s",
            "frac": dom["frac"], "traffic": ncu_traffic(dom_name), "peak_source": peak_src,
            "bytes_per_launch": dom["bytes_per_launch"], "overhead_bytes": dom["overhead_bytes"], "ms_per_launch": dom["cold_ms"],
            "shape": dom["shape"],
            "how": "dominant = largest time per step over ALL hand-written kernels; shown: its largest launch. bytes = SURVEY "
                   "§8(d) formulas (frame mix: T-frame tensor read + written once; weight gradients: x and gy read once; "
                   "lerp+concat launches with shape.act = 1 also carry the decoder block's LeakyReLU pass, reference "
                   "models/SMOW_Net.py:137: + 8*Cd*hw*s forward (z read, activated half written — it replaces BOTH the "
                   "reference's activation kernel and the copy of the decoder half), + 12*Cd*hw*s backward; with shape.act = 2 "
                   "they carry the block's BatchNorm too, :136: + 8*Cd*hw*s forward (y read, normalised + activated half written), "
                   "+ 20*Cd*hw*s backward (reduction pass + apply pass)); "
                   "time = median of CUDA-graph replays of the C-ABI call on fresh operands rotating over >= 1 GiB "
                   "(HBM-cold, no launch gaps), events on the replay stream; overhead_bytes = non-algorithmic traffic of the "
                   "same launch (e.g. the copy of the decoder half); warm = one operand set (L2-resident when it fits)",
            "all_kernels": per_kernel, "launches": sorted(rows, key=lambda r: -r["ms_per_step_cold"]),
            "hot_path_ms_per_step_cold": total, "hot_path_share_of_step": total / step_ms if step_ms > 0 else None}


def cfg3_block(args, rank, local_rank, world, device):
    """configs[2]: SMOW_Net training step, global batch 128 (strong scaling: 128/N per GPU), one CUDA graph per step."""
    import torch
    from smow_net_b200.runtime import graph as G, launch, metrics, step as S, synthetic
    gb = args.cfg3_global_batch
    lo, hi = synthetic.shard_range(gb, rank, world)
    B = hi - lo
    synthetic.seed_everything(2022, rank)
    if world > 1:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            model = launch.wrap_ddp(launch.build_model("s", device), device, world).train()
        torch.cuda.current_stream().wait_stream(side)
    else:
        model = launch.build_model("s", device).train()
    steps, warm = max(3, min(args.steps, 10)), 11 if world > 1 else 3
    opt = S.make_optimizer(model, capturable=True)
    sched = S.make_scheduler(opt, steps + 8)
    a, b, y = synthetic.make_batch(B, device=device, seed=2022 + rank)
    meter = metrics.ConfusionMeter(device)
    gs = G.GraphedStep(model, a, b, y, optimizer=opt, scheduler=sched, warmup=warm, metrics=meter)
    for _ in range(2):
        gs()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = gs()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    ms = launch.max_over_ranks(e0.elapsed_time(e1), device, world) / steps
    out = {"workload": "SMOW_Net training step (fwd + BCE-Dice + bwd + clamp-clip 0.5 + AdamW 1e-4/1e-4 + on-GPU confusion "
                       "matrix), global batch %d = %d pairs per GPU, 256x256 synthetic pairs, reference train.py:162-185" % (gb, B),
           "value": gb / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "global_batch": gb, "pairs_per_gpu": B,
           "scaling": "strong", "launch": "one CUDA graph per step (DDP all-reduce inside)" if world > 1 else "one CUDA graph per step",
           "hot_path_launches_per_step": gs.hot_path_launches, "loss": float(loss.detach()),
           "peak_mem_GB": torch.cuda.max_memory_allocated(device) / 1e9, "scores_rank0": meter.scores()}
    del gs, opt, sched, model
    return out


def cfg4_block(args, rank, world, device):
    """configs[3]: 1024x1024 tiles -> 16 crops of 256x256 each, eval, tiles sharded over the ranks, no collectives in the
    data path (the MAX of the elapsed times is taken afterwards for reporting)."""
    import torch
    from smow_net_b200.runtime import launch, synthetic
    tiles_per_gpu, per_batch = 8, 2
    total_tiles = tiles_per_gpu * world
    lo, hi = synthetic.shard_range(total_tiles, rank, world)
    model = launch.build_model("s", device).eval()
    g = torch.Generator().manual_seed(4044 + rank)
    ta = torch.randn(per_batch, 3, 1024, 1024, generator=g).pin_memory()
    tb = torch.randn(per_batch, 3, 1024, 1024, generator=g).pin_memory()

    def one_pass():
        changed = torch.zeros((), dtype=torch.int64, device=device)
        with torch.no_grad():
            for _s in range(lo, hi, per_batch):
                da, db = ta.to(device, non_blocking=True), tb.to(device, non_blocking=True)
                prob = model(synthetic.tiles_to_crops(da), synthetic.tiles_to_crops(db))
                mask = synthetic.crops_to_tiles(prob > 0.5, per_batch, 1024, 1024)        # test.py:134 threshold
                changed += mask.sum()
        return changed
    one_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    changed = one_pass()
    e1.record()
    torch.cuda.synchronize()
    ms = launch.max_over_ranks(e0.elapsed_time(e1), device, world)
    crops = total_tiles * 16
    out = {"workload": "SMOW_Net inference, %d tiles of 1024x1024 (16 crops of 256x256 each, the only size the reference "
                       "accepts), eval / no_grad, %d tiles per GPU, host tiles -> H2D inside the timed region, threshold 0.5 "
                       "(reference test.py:124-134)" % (total_tiles, tiles_per_gpu),
           "value": crops / (ms * 1e-3), "unit": UNIT, "tiles_per_s": total_tiles / (ms * 1e-3), "ms_total": ms,
           "tiles": total_tiles, "scaling": "weak", "collectives": 0, "changed_px_rank0": int(changed)}
    del model
    return out


def sweep_block(device):
    """configs[4]: isolated warp+stack sweep, HBM-cold, graph-replayed, vs ATen's own kernels on the same GPU."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "benchmarks"))
    from _aten_baseline import aten_flow_warp
    from smow_net_b200 import _lib, ops, probe
    peak, _ = measured_peak()
    clocks = ClockSampler(device.index or 0).start()
    rows = []
    for C in (64, 128, 256):
        for H in (64, 128, 256):
            per_pair = ops.warp_bwd_bytes(1, C, H, H, 4)
            B = int(max(1, min(256, -(-(1 << 30) // per_pair))))
            aten = {}
            for sigma in (0.3, 8.0):
                # ATen baseline (the reference's op sequence), fp32 NCDHW as the reference runs it; eager, CUDA events
                g = torch.Generator(device=device).manual_seed(1)
                x = torch.randn(B, C, 2, H, H, device=device, generator=g)
                flow = torch.randn(B, 2, 2, H, H, device=device, generator=g) * sigma
                gout = torch.randn(B, C, 4, H, H, device=device, generator=g)

                def t_ev(fn, n=5):
                    fn()
                    torch.cuda.synchronize()
                    ts = []
                    for _ in range(n):
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        fn()
                        b.record()
                        b.synchronize()
                        ts.append(a.elapsed_time(b))
                    return statistics.median(ts)
                with torch.no_grad():
                    f_ms = t_ev(lambda: aten_flow_warp(x, flow))
                xr, fr = x.clone().requires_grad_(True), flow.clone().requires_grad_(True)

                def fb():
                    xr.grad = fr.grad = None
                    aten_flow_warp(xr, fr).backward(gout)
                aten[sigma] = (f_ms, max(1e-6, t_ev(fb) - f_ms))
                del x, flow, gout, xr, fr
            for dt, s in ((_lib.F32, 4), (_lib.BF16, 2)):
                for lay in (_lib.NDHWC, _lib.NCDHW):
                    for sigma in (0.3, 8.0):
                        meta = {"B": B, "C": C, "H": H, "W": H, "dtype": dt, "layout": lay, "pair": 0}
                        for op, ai in (("warp_stack_fwd", 0), ("warp_stack_bwd", 1)):
                            try:
                                t = probe.time_call(op, meta, device, footprint=1 << 30, max_sets=4, sigma=sigma)
                            except RuntimeError as e:               # a (dtype, layout) combination the library refuses
                                rows.append({"op": op, "C": C, "H": H, "B": B, "dtype": "f32" if dt == _lib.F32 else "bf16",
                                             "layout": "NDHWC" if lay == _lib.NDHWC else "NCDHW", "sigma": sigma,
                                             "unsupported": str(e).splitlines()[0][:160]})
                                continue
                            rows.append({"op": op, "C": C, "H": H, "B": B, "dtype": "f32" if dt == _lib.F32 else "bf16",
                                         "layout": "NDHWC" if lay == _lib.NDHWC else "NCDHW", "sigma": sigma,
                                         "ms": t["cold_ms"], "GB_per_s": t["bytes"] / t["cold_ms"] / 1e6,
                                         "frac": t["bytes"] / t["cold_ms"] / 1e6 / peak,
                                         "aten_f32_ncdhw_ms": aten[sigma][ai], "speedup_vs_aten": aten[sigma][ai] / t["cold_ms"]})
            torch.cuda.empty_cache()
    clk = clocks.stop()

    def best(op, dtype, layout, sigma):
        r = [x["frac"] for x in rows if x["op"] == op and x["dtype"] == dtype and x["layout"] == layout and x["sigma"] == sigma
             and "frac" in x]
        return {"min_frac": min(r), "max_frac": max(r)} if r else None
    return {"workload": "BASELINE.json configs[4]: warp+stack fwd / bwd, C x (H=W) in {64,128,256}^2, fp32 + bf16, NDHWC + NCDHW, "
                        "flow sigma 0.3 (init-like) and 8 (stress); B such that one launch moves >= 1 GiB; CUDA-graph replays "
                        "(no launch gaps), operands rotating over >= 1 GiB; ATen = the reference's grid_sample + cat sequence "
                        "(fp32 NCDHW, eager, CUDA events) on the same GPU",
            "clocks": clk, "rows": rows,
            "summary": {"fwd_f32_ndhwc_s0.3": best("warp_stack_fwd", "f32", "NDHWC", 0.3),
                        "bwd_f32_ndhwc_s0.3": best("warp_stack_bwd", "f32", "NDHWC", 0.3),
                        "bwd_f32_ndhwc_s8": best("warp_stack_bwd", "f32", "NDHWC", 8.0),
                        "fwd_bf16_ndhwc_s0.3": best("warp_stack_fwd", "bf16", "NDHWC", 0.3),
                        "bwd_bf16_ndhwc_s0.3": best("warp_stack_bwd", "bf16", "NDHWC", 0.3)}}


def gpu_reference_block(args, device):
    """The like-for-like GPU baseline: the reference's own modules on the same B200, model.cuda(), eager, NCDHW, torch
    defaults (what /root/reference/train.py:122,164-179 runs)."""
    from oracle import build_ref, ref_runtime
    build_ref.build()
    if not ref_runtime.available():
        return {"unavailable": "oracle/_ref/reference.zip absent (build it with `python -m oracle.build_ref` where /root/reference exists)"}
    lw = ref_runtime.time_gpu_fwd_bwd("lw", args.batch, steps=max(3, min(args.steps, 10)), warmup=3, device=device)
    s16 = ref_runtime.time_gpu_fwd_bwd("s", 16, steps=5, warmup=2, device=device, train_step=True)
    inf = ref_runtime.time_gpu_infer("s", 32, passes=4, warmup=2, device=device)      # cfg4's batches: 2 tiles = 32 crops
    return {"kind": "reference", "how": "unmodified reference modules from oracle/_ref/reference.zip, .cuda(), eager, NCDHW, CUDA events",
            "cfg2_lw_fwd_bwd": {"value": lw["pairs_per_s"], "unit": UNIT, "ms_per_step": lw["ms_per_step"], "batch": lw["batch"]},
            "cfg3_like_s_train_step_batch16": {"value": s16["pairs_per_s"], "unit": UNIT, "ms_per_step": s16["ms_per_step"], "batch": 16},
            "cfg4_like_s_inference_32_crops": {"value": inf["pairs_per_s"], "unit": UNIT, "ms_per_pass": inf["ms_per_pass"],
                                               "crops": inf["crops"], "note": "device-resident crops, no H2D (cfg4's own number includes the H2D of its host tiles)"}}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from smow_net_b200 import _lib, ops
    from smow_net_b200.runtime import launch, step as S, synthetic
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    skip = set(x for x in args.skip.split(",") if x)
    if world > 1 and not args.no_graph:
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")      # NCCL inside a CUDA graph: no watchdog aborts
    rank, local_rank, world, device = launch.init_distributed()
    if args.strict_fp32:
        torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    synthetic.seed_everything(2022, rank)
    if world > 1 and not args.no_graph:
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):                                       # DDP must be built off the default stream
            model = launch.wrap_ddp(launch.build_model(args.model, device), device, world).train()
        torch.cuda.current_stream().wait_stream(side)
    else:
        model = launch.wrap_ddp(launch.build_model(args.model, device), device, world).train()
    B = args.batch
    a, b, y = synthetic.make_batch(B, device=device, seed=2022 + rank)
    host = [synthetic.make_batch(B, seed=3033 + rank + 7 * i, pin=True) for i in range(2)]   # two pinned batches, alternated

    # the timed step: one replay of the whole-step CUDA graph (runtime/graph.py), or the eager launch loop
    launch_mode, gs = "eager", None
    if not args.no_graph and not args.profile_step:
        try:
            from smow_net_b200.runtime import graph as G
            gs = G.GraphedStep(model, a, b, y, warmup=11 if world > 1 else max(3, args.warmup))
            launch_mode = "cuda-graph replay of the whole step (one cudaGraphLaunch per step)"
        except Exception as e:                                              # e.g. a collective that cannot be captured
            gs, launch_mode = None, "eager (graph capture failed: %s)" % (str(e).splitlines()[0][:120],)
    ok = torch.tensor([1 if gs is not None else 0], device=device)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)                           # every rank must run the same mode
        if int(ok.item()) == 0:
            gs = None
    step = (lambda: gs()) if gs is not None else (lambda: S.fwd_bwd(model, a, b, y))

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident leg (value) ----------------
    for _ in range(max(3, args.warmup)):
        step()
    sync_all()
    if args.profile_step:      # for `ncu --profile-from-start off`: exactly the timed steps are captured
        torch.cuda.profiler.start()
        for _ in range(args.steps):
            S.fwd_bwd(model, a, b, y)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print(json.dumps({"profile_step": True, "steps": args.steps}))
        return
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync_all()
    launches = gs.hot_path_launches * args.steps if gs is not None else _lib.launch_count() - l0
    ms = launch.max_over_ranks(e0.elapsed_time(e1), device, world)
    value = B * world * args.steps / (ms * 1e-3)

    # ---------------- host-buffer leg (e2e) ----------------
    # every step: H2D of that step's pinned-host batch (29.4 MB) and a D2H read of its loss.  With the graph, the copy of
    # batch n+1 is issued on the copy stream right after replay n is enqueued, so it overlaps the replay.
    def e2e_step(i):
        ha, hb, hy = host[i & 1]
        if gs is not None:
            return float(gs(ha, hb, hy, next_batch=host[(i + 1) & 1]).item())
        da, db, dy = ha.to(device, non_blocking=True), hb.to(device, non_blocking=True), hy.to(device, non_blocking=True)
        return float(S.fwd_bwd(model, da, db, dy).item())
    for i in range(3):
        e2e_step(i)
    sync_all()
    if gs is not None:
        gs._next = None                                         # the timed region starts with no batch in flight
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2.record()
    for i in range(args.steps):
        e2e_step(i)
    e3.record()
    sync_all()
    wall = time.perf_counter() - t0
    ms_e2e = launch.max_over_ranks(max(e2.elapsed_time(e3), wall * 1e3), device, world)
    clk = clocks.stop() if rank == 0 else None
    e2e_value = B * world * args.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])

    # ---------------- shapes of every hand-written launch of one real step (for the roofline block) ----------------
    with ops.kernel_timer() as kt:
        S.fwd_bwd(model, a, b, y)
        sync_all()
    calls = kt.calls()
    step_ms = ms / args.steps
    del step, gs
    model = None
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    extra = {}
    if "roofline" not in skip and rank == 0:
        roof = guarded(lambda: roofline_block(calls, step_ms, device, 1))
    else:
        roof = None
    torch.cuda.empty_cache()
    if "cfg3" not in skip:
        extra["cfg3"] = guarded(lambda: cfg3_block(args, rank, local_rank, world, device))
        gc.collect()
        torch.cuda.empty_cache()
    if "cfg4" not in skip:
        extra["cfg4"] = guarded(lambda: cfg4_block(args, rank, world, device))
        gc.collect()
        torch.cuda.empty_cache()
    if world == 1 and "sweep" not in skip:
        extra["sweep"] = guarded(lambda: sweep_block(device))
        torch.cuda.empty_cache()
    if world == 1 and "gpu_reference" not in skip:
        extra["gpu_reference"] = guarded(lambda: gpu_reference_block(args, device))
        torch.cuda.empty_cache()

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            r, kind, what = cpu_reference(args.model, B, 2, 1, 30.0)
            cb = dict({"value": r["pairs_per_s"], "unit": UNIT, "cores": r["threads"], "kind": kind,
                       "sample": "%d pairs per step x %d steps (after 1 warm-up) of %s, fwd+loss+bwd" % (r["batch"], r["steps"], what)},
                      **host_info())
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config(args, world, launch_mode), "roofline": roof, "cpu_baseline": cb,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                        "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps,
                        "how": "pinned-host batch -> staging slot on a copy stream (overlaps the previous replay) -> "
                               "device-to-device into the graph's static inputs -> replay -> loss.item()"},
                "gpu_launches": launches, "clocks": clk}
        line.update(extra)
        print(json.dumps(line))
        sys.stdout.flush()
    if world > 1:
        # captured graphs keep NCCL work objects alive and ProcessGroupNCCL's teardown then waits for ever:
        # drain the device, agree that everybody is done and leave without the teardown
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
