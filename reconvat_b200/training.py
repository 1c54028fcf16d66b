"""The caller's training step, without the host in the loop.

``train_VAT_model`` here has the signature and the arithmetic of the reference's (model/helper_functions.py:570-615)
-- ``install()`` rebinds it on request -- and differs only in what the HOST does per iteration:

* the reference calls ``loss.item()`` (:600) and formats ``sum(losses.values())`` into a progress line (:608-612) every
  iteration: two device synchronisations per optimiser step, which leave the GPU idle while Python catches up.  Here
  the running loss stays on the device and ONE line is printed per call (``log_every`` restores a progress line every
  n-th iteration);
* the VAT modules' NaN assertion (model/self_attention_VAT.py:189-190) is switched to its deferred mode for the
  duration of the loop and tested once per iteration right before ``optimizer.step()`` -- a NaN ``r_adv`` still never
  reaches the weights, but the host does not stop inside ``run_on_batch``;
* data parallelism: with ``torch.distributed`` initialised, gradients are averaged over the ranks between
  ``backward()`` and ``step()`` -- one NCCL all-reduce of the flattened gradients (2.86 M parameters = 11.4 MB for the
  ReconVAT ``UNet``: latency-bound over NVLink, a single bucket is the right granularity) -- or, for callers that want
  overlap with the backward pass, ``ddp(model)`` wraps the model so that ``run_on_batch`` runs under
  ``DistributedDataParallel``.  The hot path itself needs no collective: every rank perturbs its own segments.

``clip_grad_norm_`` keeps its place AFTER ``optimizer.step()`` (:606-607): there it only rescales gradients that the
next ``zero_grad()`` discards, i.e. it has no effect on training in the reference either.  Moving it would change the
trajectory, so it stays where it is (and launches no synchronisation).
"""
from itertools import cycle

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.nn.utils import clip_grad_norm_

__all__ = ["train_VAT_model", "allreduce_gradients", "ddp", "RunOnBatch"]


def _world(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def allreduce_gradients(model, group=None):
    """Average the gradients over the ranks with ONE all-reduce of the flattened gradient vector."""
    world = _world(group)
    if world == 1:
        return
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch._utils._flatten_dense_tensors(grads)
    dist.all_reduce(flat, group=group)
    flat.div_(world)
    torch._foreach_copy_(grads, list(torch._utils._unflatten_dense_tensors(flat, grads)))    # one call, not one per tensor


class RunOnBatch(nn.Module):
    """``forward = model.run_on_batch``: the adapter that lets ``DistributedDataParallel`` see the reference's step."""

    def __init__(self, model):
        super().__init__()
        self.model = model

    def forward(self, batch_l, batch_ul=None, VAT=False):
        return self.model.run_on_batch(batch_l, batch_ul, VAT)

    def run_on_batch(self, batch_l, batch_ul=None, VAT=False):
        return self(batch_l, batch_ul, VAT)


class _DDPRunOnBatch(nn.parallel.DistributedDataParallel):
    def run_on_batch(self, batch_l, batch_ul=None, VAT=False):
        return self(batch_l, batch_ul, VAT)


def ddp(model, **kw):
    """``DistributedDataParallel`` around ``model.run_on_batch`` (bucketed all-reduce overlapped with the backward).
    The VAT loop's inner ``autograd.grad`` only asks for the gradient of the perturbed input, so it never fires the
    reducer's parameter hooks; ``broadcast_buffers`` is off because the spectrogram tables are constants and the
    BatchNorm statistics are per rank in the sharded step."""
    kw.setdefault("broadcast_buffers", False)
    return _DDPRunOnBatch(RunOnBatch(model), **kw)


def _vat_modules(model):
    from . import VAT
    return [m for m in model.modules() if isinstance(m, VAT._VATCore)]


def train_VAT_model(model, iteration, ep, l_loader, ul_loader, optimizer, scheduler, clip_gradient_norm, alpha, VAT=False,
                    VAT_start=0, log_every=None, group=None):
    """model/helper_functions.py:570-615, same arguments and return value ``(predictions, losses, optimizer)``."""
    model.train()
    batch_size = getattr(l_loader, "batch_size", None) or 1
    wrapped = isinstance(model, nn.parallel.DistributedDataParallel)
    vats = _vat_modules(model)
    saved_strict = [v.strict for v in vats]
    for v in vats:
        v.strict = False                                     # deferred NaN assertion: tested before optimizer.step()
    total_loss = None
    l_loader = cycle(l_loader)
    if ul_loader:
        ul_loader = cycle(ul_loader)
    try:
        for i in range(iteration):
            optimizer.zero_grad()
            batch_l = next(l_loader)
            if (ep < VAT_start) or (VAT is False):
                predictions, losses, _ = model.run_on_batch(batch_l, None, False)
            else:
                batch_ul = next(ul_loader)
                predictions, losses, _ = model.run_on_batch(batch_l, batch_ul, VAT)
            loss = 0
            for key in losses.keys():
                if key.startswith('loss/train_LDS'):
                    loss += alpha * losses[key] / 2            # :591-592
                else:
                    loss += losses[key]
            loss.backward()
            total_loss = loss.detach() if total_loss is None else total_loss + loss.detach()
            if not wrapped:
                allreduce_gradients(model, group)
            for v in vats:
                v.check()                                    # AssertionError of :189-190 before the weights move
            optimizer.step()
            scheduler.step()
            if clip_gradient_norm:
                clip_grad_norm_(model.parameters(), clip_gradient_norm)      # as the reference: after the step
            if log_every and (i + 1) % log_every == 0:
                print(f'Train Epoch: {ep} [{i * batch_size}/{iteration * batch_size}'
                      f'({100. * i / iteration:.0f}%)]'
                      f"\tMain Loss: {float(sum(v.detach() for v in losses.values())):.6f}\t", end='\r')
    finally:
        for v, s in zip(vats, saved_strict):
            v.strict = s
    if total_loss is not None:
        print(' ' * 100, end='\r')
        print(f'Train Epoch: {ep}\tLoss: {float(total_loss) / iteration:.6f}')           # the one synchronisation
    return predictions, losses, optimizer
