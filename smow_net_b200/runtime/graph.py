"""Whole-step CUDA graphs: the ~2 200 kernel launches of one forward + loss + backward (+ optimiser) of
SMOW_Net / SMOW_Net_LW are captured once and replayed with a single cudaGraphLaunch, so the GPU is no
longer paced by the Python / autograd launch loop.  Streams and graphs, no tracing compiler: the kernels
are exactly the eager ones (cuDNN / ATen and the hand-written hot-path kernels of libsmow_b200.so, which
never allocate, never synchronise and launch on the current stream — capture-safe by construction).

The step semantics are runtime/step.py's (reference train.py:162-179).  Inputs live in static device
buffers.  `GraphedStep.__call__(a, b, y)` stages a new batch and replays:

* device tensors are copied into the static buffers on the replay stream;
* pinned-host tensors go through a two-slot staging ring on a separate COPY stream: the host-to-device copy of
  batch n+1 (`prefetch`) overlaps replay n, and only a device-to-device copy (HBM speed) sits between replays.
  `__call__(a, b, y, next_batch=(a2, b2, y2))` issues that prefetch right after the replay is enqueued.

Warm-up (cuDNN autotuning, lazy initialisation, DDP bucket rebuild) runs real eager steps; with an optimizer
they would move the parameters, the Adam moments and the BatchNorm statistics before the first counted step, so
every one of those tensors is snapshotted before the warm-up and restored in place after it: the first
replay starts from exactly the state the caller handed in (reference train.py:162-179 from identical weights).
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
    warmup     eager iterations on a side stream before capture (use >= 11 with DDP)
    metrics    optional runtime.metrics.ConfusionMeter: its 2x2 update (reference train.py:181-185) is captured
               in the same graph, so the per-step confusion matrix costs no host sync
    """

    def __init__(self, model, a, b, y, optimizer=None, scheduler=None, clip=0.5, warmup=3, metrics=None):
        if not a.is_cuda:
            raise RuntimeError("GraphedStep needs CUDA inputs (no CPU path)")
        self.model, self.optimizer, self.scheduler, self.clip = model, optimizer, scheduler, clip
        self.metrics = metrics
        self.a, self.b, self.y = a.clone(), b.clone(), y.clone()
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.copy_stream = torch.cuda.Stream(device=a.device)
        self._slots = [None, None]           # staging ring for pinned-host batches: [a, b, y, ready_event, free_event]
        self._next = None                    # slot index holding the prefetched batch
        self._slot_i = 0
        snapshot = self._snapshot()
        side = torch.cuda.Stream(device=a.device)
        side.wait_stream(torch.cuda.current_stream(a.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._eager()
        torch.cuda.current_stream(a.device).wait_stream(side)
        torch.cuda.synchronize(a.device)
        self._restore(snapshot)
        self._zero()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.loss = self._body()
        self.hot_path_launches = _lib.launch_count() - before    # hand-written kernels inside one replay

    # ---- state kept intact across the warm-up ---------------------------------------------------------------------
    def _snapshot(self):
        if self.optimizer is None and self.metrics is None:
            return None                       # fwd+bwd only: BatchNorm statistics move, parameters do not
        snap = {"tensors": [(t, t.detach().clone()) for t in list(self.model.parameters()) + list(self.model.buffers())]}
        return snap

    def _restore(self, snap):
        if snap is None:
            return
        with torch.no_grad():
            for t, saved in snap["tensors"]:
                t.copy_(saved)
            if self.optimizer is not None:
                for st in self.optimizer.state.values():          # AdamW's initial state: zero moments, step 0
                    for v in st.values():
                        if torch.is_tensor(v):
                            v.zero_()
            if self.metrics is not None:
                self.metrics.reset()

    def _zero(self):
        if self.optimizer is not None:
            self.optimizer.zero_grad(set_to_none=True)
        else:
            self.model.zero_grad(set_to_none=True)

    def _body(self):
        loss, pred = S.forward_loss(self.model, self.a, self.b, self.y)
        loss.backward()
        if self.optimizer is not None:
            S.clip_gradient_(self.params, self.clip)
            self.optimizer.step()
        if self.metrics is not None:
            self.metrics.update(pred.detach(), self.y)
        return loss

    def _eager(self):
        self._zero()
        return self._body()

    # ---- input staging ------------------------------------------------------------------------------------------------
    def prefetch(self, a, b, y):
        """Start the host-to-device copy of a pinned-host batch on the copy stream (overlaps the running replay)."""
        i = self._slot_i
        self._slot_i ^= 1
        slot = self._slots[i]
        if slot is None:
            slot = self._slots[i] = [torch.empty_like(self.a), torch.empty_like(self.b), torch.empty_like(self.y),
                                     torch.cuda.Event(), None]
        with torch.cuda.stream(self.copy_stream):
            if slot[4] is not None:
                self.copy_stream.wait_event(slot[4])          # the replay stream has finished reading this slot
            for dst, src in zip(slot[:3], (a, b, y)):
                dst.copy_(src, non_blocking=True)
            slot[3].record(self.copy_stream)
        self._next = i

    def load(self, a=None, b=None, y=None):
        """Copy a new batch into the static input buffers: device tensors directly; pinned-host tensors through the
        staging ring (prefetched earlier, or copied now on the copy stream)."""
        if a is None and b is None and y is None:
            return
        if a is not None and not a.is_cuda:
            if self._next is None:
                self.prefetch(a, b, y)
            slot = self._slots[self._next]
            self._next = None
            cur = torch.cuda.current_stream(self.a.device)
            cur.wait_event(slot[3])
            for dst, src in zip((self.a, self.b, self.y), slot[:3]):
                dst.copy_(src, non_blocking=True)
            slot[4] = cur.record_event()
            return
        for dst, src in ((self.a, a), (self.b, b), (self.y, y)):
            if src is not None:
                dst.copy_(src, non_blocking=True)

    def __call__(self, a=None, b=None, y=None, next_batch=None):
        self.load(a, b, y)
        self.graph.replay()
        if next_batch is not None:          # H2D of batch n+1 runs under replay n
            self.prefetch(*next_batch)
        if self.scheduler is not None:      # LR lives on the host for non-capturable schedulers: stepped outside
            self.scheduler.step()
        return self.loss
