"""Whole-model golden fixtures from the REAL reference (TEST INFRASTRUCTURE ONLY).

For each model the seeded weights of tests/helpers.seeded_model are loaded *strictly* into the
reference module (reference models/SMOW_Net.py:8-101, models/SMOW_Net_LW.py:9-99), which is then
run on the CPU on seeded inputs: an eval forward, and a train-mode forward + BCE-Dice loss +
backward (reference train.py:169-176).  Saved: sub-sampled outputs, the packed >0.5 change map,
the loss and the gradients of the OFW parameters — the tensors SURVEY §0 names as the real gates.
"""
import os
import sys

import numpy as np
import torch


def run_model(model, x1, x2, labels):
    """eval forward, then train forward/backward; returns a dict of numpy arrays."""
    out = {}
    model.eval()
    with torch.no_grad():
        y = model(x1, x2)
    out["eval_out_sub"] = y[:, :, ::4, ::4].contiguous().numpy()
    out["eval_mask_bits"] = np.packbits((y > 0.5).numpy().reshape(-1))
    model.train()
    model.zero_grad(set_to_none=True)
    y = model(x1, x2)
    pred = y.squeeze(1)
    bce = torch.nn.functional.binary_cross_entropy(pred, labels)
    dice = (2 * (labels * pred).sum() + 1e-7) / (labels.sum() + pred.sum() + 1e-7)
    loss = bce + 1 - dice
    loss.backward()
    out["train_out_sub"] = y.detach()[:, :, ::4, ::4].contiguous().numpy()
    out["train_mask_bits"] = np.packbits((y.detach() > 0.5).numpy().reshape(-1))
    out["loss"] = np.array([loss.item()], np.float64)
    sd = dict(model.named_parameters())
    for name in ("OFW.flow_make.weight", "OFW.down.0.weight", "OFW.down.6.weight", "Transformer_Encoder.conv_a.weight",
                 "decoder.conv1.weight"):
        out["grad/" + name] = sd[name].grad.detach().numpy().copy()
    out["bn/OFW.down.1.running_mean"] = dict(model.named_buffers())["OFW.down.1.running_mean"].numpy().copy()
    return out


def generate(ref_root, golden_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import helpers
    import pytest
    ref_s, ref_lw = helpers.import_reference()
    import torchvision
    import copy
    for kind in ("s", "lw"):
        mine = helpers.seeded_model(kind)
        if kind == "s":
            ref = ref_s.SMOW_Net(copy.deepcopy(torchvision.models.resnet18(weights=None)))
        else:
            ref = ref_lw.SMOW_Net_LW()
        ref.load_state_dict(copy.deepcopy(mine.state_dict()), strict=True)
        x1, x2 = helpers.seeded_pair(2)
        labels = helpers.seeded_labels(2)
        res = run_model(ref, x1, x2, labels)
        # the new module with its hot-path operators routed to the torch restatement must agree
        mp = pytest.MonkeyPatch()
        helpers.use_oracle_ops(mp)
        try:
            own = run_model(mine, x1, x2, labels)
        finally:
            mp.undo()
        for k in res:
            a, b = res[k].astype(np.float64), own[k].astype(np.float64)
            err = np.abs(a - b).max() / max(1e-12, np.abs(a).max()) if a.dtype != np.uint8 else float((res[k] != own[k]).sum())
            print("model %-2s %-40s rel.err new-module(oracle ops) vs reference: %.2e" % (kind, k, err))
            assert err < 1e-4, (kind, k, err)
        path = os.path.join(golden_dir, "model_%s.npz" % kind)
        np.savez_compressed(path, **res)
        print("wrote", path, "%.1f KiB" % (os.path.getsize(path) / 1024))
