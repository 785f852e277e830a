#!/usr/bin/env python
"""Pattern inputs for the tcgen05 weight-gradient kernel: which element lands where."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from smow_net_b200 import _lib
dev = torch.device("cuda:0"); lib = _lib.load(); CL = torch.channels_last_3d
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)

def run(x, gy, T=4, shift=1, off=1):
    B, C, _, H, W = x.shape
    x = x.contiguous(memory_format=CL); gy = gy.contiguous(memory_format=CL)
    gw = torch.full((1 + T, C, C), 7.0, device=dev)
    n = int(lib.smow_frame_mix_wgrad_tc_workspace_bytes(B, C, T, H * W))
    ws = torch.full((max(n, 16) // 4,), 5.0, dtype=torch.float32, device=dev)
    _lib.check(lib.smow_frame_mix_wgrad_tc(x.data_ptr(), gy.data_ptr(), gw.data_ptr(), B, C, T, H * W, shift, off,
                                           ws.data_ptr(), n, torch.cuda.current_stream().cuda_stream), "wgrad_tc")
    torch.cuda.synchronize()
    return gw, ws

B, C, T, H, W = 1, 32, 4, 8, 8
one = torch.ones(B, C, T, H, W, device=dev)
for dbg in (0, 16, 1):
    _lib.set_option("tc_debug", dbg)
    gw, ws = run(one, one)
    print("dbg", dbg, "ones x ones gw[0] rows 0,1,31:", gw[0, 0, :4].tolist(), gw[0, 1, :4].tolist(), gw[0, 31, 28:].tolist(), "gw[1]", gw[1, 0, :3].tolist(),
          "nonzero frac", float((gw != 0).float().mean()))
_lib.set_option("tc_debug", 0)
gw, ws = run(one, one)
print("ones x ones: expect 256 in gw[0], 64 in gw[1..4]"); print(gw[0, :4, :8]); print(gw[1, :2, :8]); print("ws head", ws[:8], "ws count of 5.0:", int((ws == 5.0).sum()), "of", ws.numel())
ci = torch.arange(1, C + 1, device=dev, dtype=torch.float32).view(1, C, 1, 1, 1)
gw, _ = run(one * ci, one)
print("x = ci+1, g = 1: expect row r = (r+1)*256"); print(gw[0, :6, :6])
gw, _ = run(one, one * ci)
print("x = 1, g = co+1: expect col c = (c+1)*256"); print(gw[0, :6, :6])
px = torch.arange(H * W, device=dev, dtype=torch.float32).view(1, 1, 1, H, W)
gw, _ = run(one * px, one)
print("x = pixel idx, g = 1: expect %g" % (4 * float(px.sum()))); print(gw[0, :3, :6])
fr = torch.arange(1, T + 1, device=dev, dtype=torch.float32).view(1, 1, T, 1, 1)
gw, _ = run(one * fr, one)
print("x = frame+1, g=1: gw[0] expect 64*(1+2+3+4)=640; gw[1+g]: f=(g-1)%4, src=(f+1)%4=g -> 64*(g+1)"); print(gw[:, 0, :4])
