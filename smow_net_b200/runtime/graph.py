"""Whole-step CUDA graphs: the ~2 200 kernel launches of one forward + loss + backward (+ optimiser) of
SMOW_Net / SMOW_Net_LW are captured once and replayed with a single cudaGraphLaunch, so the GPU is no
longer paced by the Python / autograd launch loop.  Streams and graphs, no tracing compiler: the kernels
are exactly the eager ones (cuDNN / ATen and the hand-written hot-path kernels of libsmow_b200.so, which
never allocate, never synchronise and launch on the current stream — capture-safe by construction).

The step semantics are runtime/step.py's (reference train.py:162-179).  Inputs live in static device
buffers; `GraphedStep.__call__(a, b, y)` copies new inputs in (device or pinned-host tensors) and replays.
"""
import torch

from .. import _lib
from . import step as S


class GraphedStep:
    """fwd_bwd (optimizer is None) or the full train_step captured in one CUDA graph.

    model      the (optionally DDP-wrapped) network, in the mode it is to be run in
    a, b, y    example inputs (B,3,256,256) x2 and labels (B,256,256); their shapes are frozen
    optimizer  capture clip + AdamW step too (must be capturable: torch.optim.AdamW(..., capturable=True)
               or fused=True on CUDA parameters)
    warmup     eager iterations on a side stream before capture (cuDNN autotuning, lazy initialisation,
               DDP bucket rebuild: use >= 11 with DDP)
    """

    def __init__(self, model, a, b, y, optimizer=None, scheduler=None, clip=0.5, warmup=3):
        if not a.is_cuda:
            raise RuntimeError("GraphedStep needs CUDA inputs (no CPU path)")
        self.model, self.optimizer, self.scheduler, self.clip = model, optimizer, scheduler, clip
        self.a, self.b, self.y = a.clone(), b.clone(), y.clone()
        self.params = [p for p in model.parameters() if p.requires_grad]
        side = torch.cuda.Stream(device=a.device)
        side.wait_stream(torch.cuda.current_stream(a.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream(a.device).wait_stream(side)
        torch.cuda.synchronize(a.device)
        self._zero()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        self.hot_path_launches = _lib.launch_count() - before    # hand-written kernels inside one replay

    def _zero(self):
        if self.optimizer is not None:
            self.optimizer.zero_grad(set_to_none=True)
        else:
            self.model.zero_grad(set_to_none=True)

    def _body(self):
        loss, _ = S.forward_loss(self.model, self.a, self.b, self.y)
        loss.backward()
        if self.optimizer is not None:
            S.clip_gradient_(self.params, self.clip)
            self.optimizer.step()
        return loss

    def _eager(self):
        self._zero()
        return self._body()

    def load(self, a=None, b=None, y=None):
        """Copy a new batch into the static input buffers (asynchronous for pinned host tensors)."""
        for dst, src in ((self.a, a), (self.b, b), (self.y, y)):
            if src is not None:
                dst.copy_(src, non_blocking=True)

    def __call__(self, a=None, b=None, y=None):
        self.load(a, b, y)
        self.graph.replay()
        if self.scheduler is not None:      # LR lives on the host for non-capturable schedulers: stepped outside
            self.scheduler.step()
        return self.loss
