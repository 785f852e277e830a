"""CPU port of the reference's whole model for bench.py's cpu_baseline / --impl reference legs.
TEST / MEASUREMENT INFRASTRUCTURE ONLY.

The reference tree cannot travel to the GPU box (it is Python, not installable: no setup.py /
pyproject), so its CPU implementation is reproduced as: the drop-in module's cuDNN-free plumbing
(bit-identical to the reference's on the CPU, see tests/test_models_cpu.py) + the torch restatement of the
hot-path operators (oracle/torch_ref.py), i.e. exactly the ATen op sequence the reference executes."""
import contextlib

from . import torch_ref


@contextlib.contextmanager
def reference_ops():
    """Temporarily route smow_net_b200.ops' four operators to the reference's ATen op sequence."""
    from smow_net_b200 import ops
    saved = {k: getattr(ops, k) for k in ("flow_warp", "tlerp", "tlerp_cat", "tlerp_pair_cat")}
    ops.flow_warp = lambda x, flow, size=None: torch_ref.ref_flow_warp(x, flow)
    ops.tlerp = torch_ref.ref_tlerp
    ops.tlerp_cat = torch_ref.ref_tlerp_cat
    ops.tlerp_pair_cat = lambda dec, a, b: torch_ref.ref_tlerp_cat(dec, torch_ref.ref_pair_stack(a, b))
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(ops, k, v)


def time_cpu_fwd_bwd(kind, batch, steps, warmup, budget_s=150.0, threads=None):
    """pairs/s of the reference-port model's fwd+loss+bwd on the host cores.  The per-step sample is
    `batch` pairs, shrunk (to >= 1) when the first step shows the run would exceed `budget_s`."""
    import os
    import time
    import torch
    from smow_net_b200.runtime import step as S
    from smow_net_b200.runtime import synthetic
    from smow_net_b200.runtime.launch import build_model
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    with reference_ops():
        torch.manual_seed(2022)
        model = build_model(kind, torch.device("cpu")).train()
        for prm in model.parameters():          # the reference keeps its weights in the default (contiguous) order
            if prm.dim() >= 4:
                prm.data = prm.data.contiguous()
        a, b, y = synthetic.make_batch(batch)
        t0 = time.perf_counter()
        S.fwd_bwd(model, a, b, y)                     # first (untimed) step doubles as the probe
        probe = time.perf_counter() - t0
        n = batch
        total = max(1, steps + max(0, warmup - 1))
        if probe * total > budget_s:
            n = max(1, int(batch * budget_s / (probe * total)))
            a, b, y = a[:n].contiguous(), b[:n].contiguous(), y[:n].contiguous()
        for _ in range(max(0, warmup - 1)):
            S.fwd_bwd(model, a, b, y)
        t0 = time.perf_counter()
        for _ in range(steps):
            S.fwd_bwd(model, a, b, y)
        dt = time.perf_counter() - t0
    return {"pairs_per_s": n * steps / dt, "ms_per_step": 1e3 * dt / steps, "batch": n, "threads": threads,
            "steps": steps}
