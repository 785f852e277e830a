#!/usr/bin/env python
"""bench.py — image-pairs/s of SMOW_Net_LW fwd+bwd (256x256, batch 16 per GPU) + hot-path roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Contract: W untimed warm-up steps, exactly K timed steps bracketed by
barrier + synchronize, CUDA-event timing, MAX over ranks, one JSON line from rank 0.
  value     whole-job pairs/s with inputs resident in HBM
  e2e       same metric through the public module call with pinned HOST inputs (H2D inside the timed
            region) and a D2H read of the loss every step
  roofline  dominant hand-written kernel: algorithmic bytes / CUDA-event duration vs measured HBM peak
  cpu_baseline   the oracle's CPU port of the reference model on the host cores (bounded sample)
--impl reference times that CPU port alone (the reference is pure Python and cannot be installed on
the box; oracle/cpu_model.py explains what is timed).
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
    ap.add_argument("--strict-fp32", action="store_true", help="disable cuDNN TF32 convolutions")
    ap.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying its CUDA graph")
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
            "conv_math": "fp32 hot path; cuDNN convolutions under torch defaults (allow_tf32=%s)"
                         % (not args.strict_fp32)}


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
        return json.load(open(p)).get(kernel)
    return None


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import cpu_model
    r = cpu_model.time_cpu_fwd_bwd(args.model, args.batch, args.steps, args.warmup)
    cb = {"value": r["pairs_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
          "sample": "%d pairs per step x %d steps, fwd+loss+bwd of the reference-port model on the host CPU"
                    % (r["batch"], r["steps"])}
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": r["pairs_per_s"], "unit": UNIT,
                      "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": config(args, 1), "cpu_baseline": cb,
                      "e2e": {"value": r["pairs_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                      "gpu_launches": 0}))


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
    ha, hb, hy = synthetic.make_batch(B, seed=3033 + rank, pin=True)

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
    def e2e_step():
        if gs is not None:                                     # H2D into the graph's static input buffers, replay
            return float(gs(ha, hb, hy).item())                # D2H read of the loss every step
        da, db, dy = ha.to(device, non_blocking=True), hb.to(device, non_blocking=True), hy.to(device, non_blocking=True)
        return float(S.fwd_bwd(model, da, db, dy).item())
    for _ in range(3):
        e2e_step()
    sync_all()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e2.record()
    for _ in range(args.steps):
        e2e_step()
    e3.record()
    sync_all()
    wall = time.perf_counter() - t0
    ms_e2e = launch.max_over_ranks(max(e2.elapsed_time(e3), wall * 1e3), device, world)
    clk = clocks.stop() if rank == 0 else None
    e2e_value = B * world * args.steps / (ms_e2e * 1e-3)
    # per-launch CUDA-event timing of the hand-written kernels inside real (eager) steps, for the roofline object:
    # events cannot be recorded inside a graph replay, so these steps run outside the timed regions
    n_kt = min(5, args.steps)
    with ops.kernel_timer() as kt:
        for _ in range(n_kt):
            S.fwd_bwd(model, a, b, y)
        sync_all()
    h2d = sum(t.numel() * t.element_size() for t in (ha, hb, hy))

    if rank == 0:
        peak, peak_src = measured_peak()
        summ, shapes = kt.summary(), kt.summary(by_shape=True)
        kernels = {k: {"calls_per_step": v["calls"] / n_kt, "ms_per_call": v["ms"] / v["calls"],
                       "algorithmic_GB_per_s": v["gbps"], "frac_of_peak": v["gbps"] / peak,
                       "ms_per_step": v["ms"] / n_kt,
                       "bound": "fp32 issue (~300 FP32 instructions per 64-byte pixel), not HBM" if k.startswith("tokenizer")
                                else "hbm", "row": "N2" if k.startswith("tokenizer") else "N4" if k.startswith("frame_mix")
                                else "A1-A5"} for k, v in summ.items()}
        # the dominant LAUNCH of the path BASELINE.json names (warp+stack and temporal
        # lerp+concat): launches of one operator differ by 300x in size across the decoder levels, so they are kept
        # apart by shape; the widened rows (N2 tokenizer: FP32-issue bound; N4 frame mix) are reported in all_kernels only
        hot = {k: v for k, v in shapes.items() if k.startswith(("warp_", "tlerp_"))}      # rows N2 / N4: all_kernels only
        dom = max(hot, key=lambda k: hot[k]["ms"])
        roof = {"bound": "hbm", "kernel": dom.split("@")[0], "achieved": hot[dom]["gbps"], "peak": peak, "unit": "GB/s",
                "frac": hot[dom]["gbps"] / peak, "traffic": ncu_traffic(dom.split("@")[0] + "@largest"), "peak_source": peak_src,
                "bytes_per_launch": hot[dom]["bytes"] / hot[dom]["calls"],
                "ms_per_launch": hot[dom]["ms"] / hot[dom]["calls"],
                "note": "CUDA events around each C-ABI call inside %d real (eager) steps after the timed regions; "
                        "in-step launches: operands were just produced, so part of the traffic is L2-resident and the "
                        "events include launch latency; HBM-cold figures are in profiles/ (benchmarks/sweep_warp.py)" % n_kt,
                "all_kernels": kernels,
                "hot_path_share_of_step": sum(v["ms"] for v in summ.values()) / n_kt / (ms / args.steps if ms > 0 else 1)}
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_model
            r = cpu_model.time_cpu_fwd_bwd(args.model, B, steps=2, warmup=1, budget_s=30.0)
            cb = {"value": r["pairs_per_s"], "unit": UNIT, "cores": r["threads"], "kind": "port",
                  "sample": "%d pairs per step x %d steps (after 1 warm-up) of the reference-port model, fwd+loss+bwd"
                            % (r["batch"], r["steps"])}
        print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                          "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config(args, world, launch_mode), "roofline": roof, "cpu_baseline": cb,
                          "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                                  "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
                          "gpu_launches": launches, "clocks": clk}))
    if world > 1:
        if gs is not None:
            # a captured graph keeps NCCL work objects alive and ProcessGroupNCCL's teardown then waits for ever:
            # drop the graph, drain the device, agree that everybody is done and leave without the teardown
            del step, gs
            import gc
            gc.collect()
            torch.cuda.synchronize()
            dist.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
