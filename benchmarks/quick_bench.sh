#!/bin/bash
# headline + roofline only (no cfg3 / cfg4 / sweep / reference legs): the A/B loop while tuning a kernel
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 --skip cfg3,cfg4,sweep,gpu_reference > gpurun_out/quick.json 2> gpurun_out/quick.err || tail -5 gpurun_out/quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/quick.json").read().strip().splitlines()[-1])
print("value %.1f pairs/s  %.3f ms/step  e2e %.1f  clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"]))
r = d["roofline"]
print("dominant", r["kernel"], "%.3f" % r["frac"], "%.1f us" % (r["ms_per_launch"] * 1e3), r["shape"])
for k, v in r["all_kernels"].items():
    print("  %-16s %.3f ms/step cold  %.3f warm  frac %.2f" % (k, v["ms_per_step"], v["ms_per_step_warm"], v["frac"]))
PY
