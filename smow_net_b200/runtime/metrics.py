"""On-GPU change-detection metrics (row N3 of SURVEY §8f).

The reference copies every step's predictions to the host and builds a 2x2 confusion matrix with numpy
(train.py:181-185, utils/metric_tool.py:132-155) — a device-to-host sync per step that serialises every GPU.
`ConfusionMeter` keeps the same matrix as four int64 counters on the device: the update is four reductions with
no host interaction (capturable in the step's CUDA graph), and the scores are read back once per epoch.
"""
import torch

_EPS = float(torch.finfo(torch.float32).eps)      # np.finfo(np.float32).eps in utils/metric_tool.py


class ConfusionMeter:
    """cm[gt][pred] over {0: unchanged, 1: changed}; `update` thresholds at 0.5 like train.py:182."""

    def __init__(self, device):
        self.cm = torch.zeros(2, 2, dtype=torch.int64, device=device)

    def reset(self):
        self.cm.zero_()

    @torch.no_grad()
    def update(self, pred, gts):
        """pred: probabilities (B,H,W); gts: {0,1} labels (B,H,W).  No host sync, no data-dependent shapes."""
        p = pred > 0.5
        g = gts.to(torch.int64) == 1                      # mask = gts.numpy().astype(int)  (train.py:183)
        n = p.numel()
        sp, sg, tp = p.sum(dtype=torch.int64), g.sum(dtype=torch.int64), (p & g).sum(dtype=torch.int64)
        self.cm += torch.stack((n - sp - sg + tp, sp - tp, sg - tp, tp)).view(2, 2)

    def all_reduce(self):
        """Optional end-of-epoch reduction of the counters over the data-parallel ranks."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.cm)

    def scores(self):
        """acc / iou / F1 / precision / recall of the 'changed' class — utils/metric_tool.py:93-129 (cm2score).
        ONE device-to-host read."""
        h = self.cm.to("cpu", torch.float64)
        tp, total = float(h[1, 1]), float(h.sum())
        tpfn, tpfp = float(h[1].sum()), float(h[:, 1].sum())
        acc = float(h.diagonal().sum()) / (total + _EPS)
        recall = tp / (tpfn + _EPS)
        precision = tp / (tpfp + _EPS)
        f1 = 2 * ((recall * precision) / (recall + precision + _EPS))
        iou = tp / (total - float(h[0, 0]) + _EPS)
        return {"acc": acc, "iou": iou, "F1": f1, "precision": precision, "recall": recall}
