"""The reference's optimisation step, restated (train.py:162-179 of the reference):
zero_grad -> forward -> squeeze(1) -> BCE + (1 - dice) -> backward -> clamp-clip +-0.5 -> AdamW(1e-4, wd 1e-4)
-> per-iteration cosine LR.  The per-step device-to-host metric copy (train.py:181-185) is NOT part of the
step here: it is a host sync that would serialise every GPU."""
import torch
import torch.nn.functional as F


def bce_dice_loss(pred, true, eps=1e-7):
    """utils/loss_f.py:8-18: BCELoss(pred, true) + 1 - global dice."""
    bce = F.binary_cross_entropy(pred, true)          # first, like the reference: autograd then sums the three gradient
    inter = (true * pred).sum()                       # contributions to `pred` in the same order (bit-identical grads)
    return bce + 1 - (2 * inter + eps) / (true.sum() + pred.sum() + eps)


def clip_gradient_(params, clip):
    """utils/func.py:4-8: element-wise clamp of every gradient (not a norm clip), as one fused
    multi-tensor call instead of the reference's ~280 clamp_ launches."""
    grads = [p.grad for p in params if p.grad is not None]
    if grads:
        torch._foreach_clamp_min_(grads, -clip)
        torch._foreach_clamp_max_(grads, clip)


def make_optimizer(model, lr=1e-4, weight_decay=1e-4, capturable=False):
    """AdamW(1e-4, wd 1e-4) as train.py:127; fused multi-tensor kernel on CUDA.  capturable=True keeps the step count
    and the learning rate on the device so the step can live inside a CUDA graph (runtime/graph.py) while the
    scheduler keeps updating the rate from the host."""
    params = [p for p in model.parameters() if p.requires_grad]
    fused = bool(params) and all(p.is_cuda for p in params)    # one multi-tensor kernel instead of ~10 per tensor
    if capturable and fused:
        lr = torch.tensor(float(lr), device=params[0].device)
        return torch.optim.AdamW(params, lr, weight_decay=weight_decay, fused=True, capturable=True)
    return torch.optim.AdamW(params, lr, weight_decay=weight_decay, fused=fused)


def make_scheduler(optimizer, total_iters):
    """utils/lr_scheduler.py:64-69: CosineAnnealingLR stepped per iteration, eta_min 1e-6."""
    return torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=max(1, total_iters), eta_min=1e-6)


def forward_loss(model, image_a, image_b, labels):
    pred = model(image_a, image_b)
    pred = pred.squeeze(1) if pred.size(1) == 1 else torch.sigmoid(pred)[:, 1]
    return bce_dice_loss(pred, labels), pred


def fwd_bwd(model, image_a, image_b, labels):
    """BASELINE.json's 'fwd+bwd' unit of work: zero_grad, forward, loss, backward (gradient all-reduce
    included when the model is wrapped in DDP)."""
    model.zero_grad(set_to_none=True)
    loss, _ = forward_loss(model, image_a, image_b, labels)
    loss.backward()
    return loss


def train_step(model, optimizer, scheduler, image_a, image_b, labels, clip=0.5):
    optimizer.zero_grad(set_to_none=True)
    loss, _ = forward_loss(model, image_a, image_b, labels)
    loss.backward()
    clip_gradient_([p for g in optimizer.param_groups for p in g["params"]], clip)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return loss


def freeze_unused(model):
    """SMOW_Net_LW runs backbone.features[18] and discards the result (reference
    models/SMOW_Net_LW.py:573-579), so its 3 parameter tensors never receive a gradient; DDP needs them
    excluded (or find_unused_parameters=True, which costs a graph walk per step)."""
    feats = getattr(getattr(model, "backbone", None), "features", None)
    if feats is not None:
        for p in feats[18].parameters():
            p.requires_grad_(False)
    return model
