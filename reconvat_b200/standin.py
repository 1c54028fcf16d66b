"""Tiny deterministic stand-in for the transcription network.

The VAT loop treats the network as a black-box callable (it calls
``model.transcriber(x)`` or ``model(x)`` and unpacks a tuple --
model/self_attention_VAT.py:164, model/UNet_onset.py:118,
model/onset_frame_VAT.py:177).  The real U-Net is out of scope (SURVEY.md
section 8), so tests, golden fixtures and the benchmark drive the loop with
this frame-wise ``sigmoid(x @ W + b)`` whose weights come from an integer
hash (no framework RNG), in the four call conventions the reference uses.
"""
import numpy as np
import torch
import torch.nn as nn

from . import synth


def _hash_normal(n, seed):
    u1 = synth.uniform01(n, seed)
    u2 = synth.uniform01(n, seed + 7777)
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2 * np.pi * u2)


class _Head(nn.Module):
    def __init__(self, n_in, n_out, seed, gain):
        super().__init__()
        w = _hash_normal(n_in * n_out, seed).reshape(n_in, n_out) * gain / np.sqrt(n_in)
        b = _hash_normal(n_out, seed + 1) * 0.5
        self.weight = nn.Parameter(torch.tensor(w, dtype=torch.float32))
        self.bias = nn.Parameter(torch.tensor(b, dtype=torch.float32))

    def forward(self, x):
        if x.dim() == 4:                      # (B,1,T,F) -> (B,T,F)
            x = x.squeeze(1)
        return torch.sigmoid((x - 0.5) @ self.weight + self.bias)


class StandInTranscriber(nn.Module):
    """convention:
    'unet'        transcriber(x) -> (frame, None)            self_attention_VAT.UNet
    'unet_onset'  transcriber(x) -> (frame, onset, None)     UNet_onset.UNet_Onset
    'stepwise'    model(x)       -> (frame, None)            self_attention_VAT.stepwise_VAT / VAT.py
    'onf'         model(x)       -> (onset, activation, frame)   onset_frame_VAT (x is 3-D)
    'stack'       model(x)       -> (activation, frame)      onset_frame_VAT.stepwise_VAT_frame_stack
    'seg'         model(x)       -> frame                    Segmentation.Seg_VAT (the posterior itself)
    """

    def __init__(self, convention="unet", n_in=229, n_out=88, seed=0, gain=6.0):
        super().__init__()
        self.convention = convention
        self.frame = _Head(n_in, n_out, 100 + seed, gain)
        self.onset = _Head(n_in, n_out, 200 + seed, gain) if convention in ("unet_onset", "onf", "stack") else None
        self.captured_grads = None            # set to a list to record dL/dx_adv of grad-requiring inputs
        self.transcriber = self._transcriber if convention in ("unet", "unet_onset") else None

    def _maybe_capture(self, x):
        if self.captured_grads is not None and x.requires_grad:
            x.register_hook(lambda g: self.captured_grads.append(g.detach().clone()))

    def _transcriber(self, x):
        self._maybe_capture(x)
        if self.convention == "unet":
            return self.frame(x), None
        return self.frame(x), self.onset(x), None

    def forward(self, x):
        self._maybe_capture(x)
        if self.convention == "stepwise":
            return self.frame(x), None
        if self.convention == "onf":
            f = self.frame(x)
            o = self.onset(x)
            return o, 0.5 * (o + f), f
        if self.convention == "stack":
            f = self.frame(x)
            return 0.5 * (self.onset(x) + f), f
        if self.convention == "seg":
            return self.frame(x)
        raise RuntimeError("call .transcriber(x) for convention %r" % self.convention)


class _InjectGrad(torch.autograd.Function):
    """forward: hand back a precomputed posterior; backward: hand back a precomputed dL/dx_adv."""

    @staticmethod
    def forward(ctx, x, y, g):
        ctx.save_for_backward(g)
        return y.view_as(y)

    @staticmethod
    def backward(ctx, grad_y):
        (g,) = ctx.saved_tensors
        return g, None, None


class InjectedTranscriber(nn.Module):
    """A zero-cost 'network' for measuring the hot path alone (headline step: "one VAT
    call's kernels with injected g"): ``transcriber(x)`` cycles through three precomputed posteriors
    (clean, XI-perturbed, eps-perturbed) without launching a kernel, and the backward of the second one
    returns the precomputed input gradient ``g``.  Works on CPU and CUDA, so the reference arm of the
    benchmark runs the same workload."""

    def __init__(self, batch, frames=640, mels=229, pitches=88, seed=0, channel_dim=True):
        super().__init__()
        n = batch * frames * pitches
        ys = [1.0 / (1.0 + np.exp(-3.0 * _hash_normal(n, seed + 10 * i).reshape(batch, frames, pitches)))
              for i in range(3)]
        for i, y in enumerate(ys):
            self.register_buffer("y%d" % i, torch.tensor(y, dtype=torch.float32))
        shape = (batch, 1, frames, mels) if channel_dim else (batch, frames, mels)
        self.register_buffer("g", torch.tensor(_hash_normal(batch * frames * mels, seed + 99).reshape(shape) * 1e-7,
                                               dtype=torch.float32))
        self._call = 0

    def transcriber(self, x):
        i = self._call % 3
        self._call += 1
        if i == 1 and x.requires_grad:
            return _InjectGrad.apply(x, self.y1, self.g), None
        return getattr(self, "y%d" % i), None

    def forward(self, x):
        return self.transcriber(x)
