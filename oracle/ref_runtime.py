"""The REAL reference modules, loaded from oracle/_ref/reference.zip (see oracle/build_ref.py) or, in the build
container, straight from /root/reference — TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Used by bench.py's `--impl reference` arm, its `cpu_baseline` leg and its `gpu_reference` leg (the reference's own
module, `model.cuda()`, eager, NCDHW — what train.py:122,164-179 runs), and by tests that pin runtime/step.py against
the reference's utils/func.py and utils/loss_f.py.  The product package never imports this file.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ZIP = os.path.join(HERE, "_ref", "reference.zip")
_mods = None


def available():
    return os.path.exists(ZIP)


def modules():
    """dict(SMOW_Net, SMOW_Net_LW, func, loss_f, metric_tool) of the unmodified reference modules, or None."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        return None
    if ZIP not in sys.path:
        sys.path.insert(0, ZIP)
    m = {"SMOW_Net": importlib.import_module("smow_reference.models.SMOW_Net"),
         "SMOW_Net_LW": importlib.import_module("smow_reference.models.SMOW_Net_LW"),
         "func": importlib.import_module("smow_reference.utils.func"),
         "loss_f": importlib.import_module("smow_reference.utils.loss_f"),
         "metric_tool": importlib.import_module("smow_reference.utils.metric_tool")}
    # SMOW_Net_LW() downloads ImageNet weights in its constructor (models/SMOW_Net_LW.py:12,584); there is no network
    # and BASELINE.json asks for random-init weights: an empty state_dict keeps the random initialisation
    m["SMOW_Net_LW"].load_state_dict_from_url = lambda *a, **k: {}
    _mods = m
    return m


def build_model(kind, device="cpu", seed=2022):
    """The reference's own model object: SMOW_Net(deepcopy(resnet18(weights=None))) (train.py:84-85) or SMOW_Net_LW()."""
    import copy
    import torch
    m = modules()
    if m is None:
        raise RuntimeError("oracle/_ref/reference.zip is absent: run `python -m oracle.build_ref` in the build container")
    import contextlib
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(sys.stderr):          # the LW constructor prints; bench.py's stdout is one JSON line
        if kind == "s":
            import torchvision
            model = m["SMOW_Net"].SMOW_Net(copy.deepcopy(torchvision.models.resnet18(weights=None)))
        else:
            model = m["SMOW_Net_LW"].SMOW_Net_LW()
    return model.to(device)


def fwd_bwd(model, a, b, y, loss_fn):
    """train.py:162-176 without the optimiser: zero_grad, forward, squeeze(1), BCEDICE_loss, backward."""
    model.zero_grad(set_to_none=True)
    pred = model(a, b)
    pred = pred.squeeze(1) if pred.size(1) == 1 else __import__("torch").sigmoid(pred)[:, 1]
    loss = loss_fn(pred, y)
    loss.backward()
    return loss


def time_cpu_fwd_bwd(kind, batch, steps, warmup, budget_s=150.0, threads=None):
    """pairs/s of the reference model's fwd+loss+bwd on the host cores (all of them unless `threads` is given).
    The per-step sample is `batch` pairs, shrunk (to >= 1) when the first step shows the run would exceed `budget_s`."""
    import time
    import torch
    from smow_net_b200.runtime import synthetic
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    model = build_model(kind).train()
    loss_fn = modules()["loss_f"].BCEDICE_loss
    a, b, y = synthetic.make_batch(batch)
    t0 = time.perf_counter()
    fwd_bwd(model, a, b, y, loss_fn)                      # first (untimed) step doubles as the probe
    probe = time.perf_counter() - t0
    n = batch
    total = max(1, steps + max(0, warmup - 1))
    if probe * total > budget_s:
        n = max(1, int(batch * budget_s / (probe * total)))
        a, b, y = a[:n].contiguous(), b[:n].contiguous(), y[:n].contiguous()
    for _ in range(max(0, warmup - 1)):
        fwd_bwd(model, a, b, y, loss_fn)
    t0 = time.perf_counter()
    for _ in range(steps):
        fwd_bwd(model, a, b, y, loss_fn)
    dt = time.perf_counter() - t0
    return {"pairs_per_s": n * steps / dt, "ms_per_step": 1e3 * dt / steps, "batch": n, "threads": threads,
            "steps": steps, "kind": "reference"}


def time_gpu_fwd_bwd(kind, batch, steps, warmup, device, train_step=False):
    """The reference on the same GPU the way train.py runs it: model.cuda(), eager, NCDHW, torch defaults.
    train_step=True adds clip_gradient + AdamW(1e-4, wd 1e-4) (train.py:135,177-178).  CUDA events."""
    import torch
    from smow_net_b200.runtime import synthetic
    m = modules()
    model = build_model(kind, device).train()
    loss_fn = m["loss_f"].BCEDICE_loss
    a, b, y = synthetic.make_batch(batch, device=device)
    opt = torch.optim.AdamW(model.parameters(), 1e-4, weight_decay=1e-4) if train_step else None

    def step():
        loss = fwd_bwd(model, a, b, y, loss_fn)
        if opt is not None:
            m["func"].clip_gradient(opt, 0.5)
            opt.step()
        return loss

    for _ in range(max(1, warmup)):
        step()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
    del model, opt
    torch.cuda.empty_cache()
    return {"pairs_per_s": batch * steps / (ms * 1e-3), "ms_per_step": ms / steps, "batch": batch, "steps": steps}


def time_gpu_infer(kind, crops, passes, warmup, device):
    """The reference's inference the way test.py:124-134 runs it on the same GPU: model.cuda().eval(), no_grad, eager, NCDHW,
    batches of `crops` 256x256 crops (one 1024x1024 tile = 16 crops), threshold 0.5.  CUDA events."""
    import torch
    from smow_net_b200.runtime import synthetic
    model = build_model(kind, device).eval()
    a, b, _ = synthetic.make_batch(crops, device=device)

    def one():
        with torch.no_grad():
            return (model(a, b) > 0.5).sum()

    for _ in range(max(1, warmup)):
        one()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(passes):
        one()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1)
    del model
    torch.cuda.empty_cache()
    return {"pairs_per_s": crops * passes / (ms * 1e-3), "ms_per_pass": ms / passes, "crops": crops, "passes": passes}
