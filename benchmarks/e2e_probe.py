#!/usr/bin/env python
"""Step-time experiments around the (cuDNN/ATen-bound) rest of the SMOW_Net_LW step: which library paths are slow?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from smow_net_b200.runtime import launch, step as S, synthetic  # noqa: E402


def measure(model, a, b, y, steps=10):
    for _ in range(4):
        S.fwd_bwd(model, a, b, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        S.fwd_bwd(model, a, b, y)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    kind = sys.argv[1] if len(sys.argv) > 1 else "lw"
    a, b, y = synthetic.make_batch(16, device=dev)
    torch.manual_seed(0)
    model = launch.build_model(kind, dev).train()
    print("baseline            %.2f ms" % measure(model, a, b, y))
    orig = F.batch_norm

    def native_bn(*args, **kw):
        with torch.backends.cudnn.flags(enabled=False):
            return orig(*args, **kw)
    F.batch_norm = native_bn
    torch.nn.functional.batch_norm = native_bn
    print("native batch_norm   %.2f ms" % measure(model, a, b, y))
    F.batch_norm = orig
    if kind == "lw":
        model.backbone.to(memory_format=torch.channels_last)
        a2, b2 = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)
        print("backbone NHWC       %.2f ms" % measure(model, a2, b2, y))
        for name, p in model.named_parameters():                # 5-D parameters: decoder / OFW convs
            if p.dim() == 5:
                p.data = p.data.contiguous(memory_format=torch.channels_last_3d)
        try:
            print("+ decoder NDHWC     %.2f ms" % measure(model, a2, b2, y))
        except Exception as e:                                   # noqa: BLE001
            print("decoder NDHWC failed:", repr(e)[:300])
        torch.backends.cudnn.benchmark = False
        print("  cudnn.benchmark=0 %.2f ms" % measure(model, a2, b2, y))
    else:
        for name, p in model.named_parameters():
            if p.dim() == 5:
                p.data = p.data.contiguous(memory_format=torch.channels_last_3d)
        try:
            print("all 5-D NDHWC       %.2f ms" % measure(model, a, b, y))
        except Exception as e:                                   # noqa: BLE001
            print("NDHWC failed:", repr(e)[:300])


if __name__ == "__main__":
    main()
