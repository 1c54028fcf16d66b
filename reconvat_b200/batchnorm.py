"""``nn.BatchNorm2d`` of the caller's U-Net on hand-written kernels (consumer side of the hot path).

The reference puts an ``nn.BatchNorm2d(C)`` with C = 16 .. 128 behind every convolution of its two U-Nets
(model/self_attention_VAT.py:848-850, :865-869; model/UNet_onset.py:190-211): 30 modules, three forward and two backward
passes per training iteration.  cuDNN runs that layout with one block per channel; on a B200 the two kernels take half of
the iteration (profiles/r02_train_step.md).  :class:`BatchNorm2d` is the same module -- same constructor, parameters,
buffers (``state_dict`` interchange), train / eval semantics, running-statistics update -- over ``rvb_bn_*``: every
channel cut into slices so that the whole chip works on it, two launches per direction, float64 partial sums.
Both memory formats have kernels of their own: NCHW (what the reference's unchanged scripts produce) and
``torch.channels_last`` (C % 4 == 0).  For channels_last input the MODULE hands the call to cuDNN by default -- its NHWC
kernels are sound and ATen's host cost per call is lower -- and runs the rvb NHWC kernels with ``RVB_BN_NHWC=1``; it
never forces a layout conversion on its neighbours either way.

``convert(model)`` swaps the class of every ``nn.BatchNorm2d`` in place (parameters untouched);
``install(batchnorm=True)`` makes the reference's model files construct this class without editing them.
There is no CPU path: a CPU tensor raises.
"""
import os

import torch
import torch.nn as nn

from . import _lib

__all__ = ["BatchNorm2d", "convert", "batch_norm"]


def _ptr(t):
    return None if t is None else t.data_ptr()


_WORKSPACE = {}
_MAX_SPLITS = 64


def _partials(device, c, stream):
    """float64 [c][64][2] workspace of the two reductions, one per (device, stream): calls on one stream are ordered.
    Under CUDA-graph capture every call gets its own (graphs captured on one stream may be replayed side by side)."""
    if torch.cuda.is_current_stream_capturing():
        return torch.empty((c * _MAX_SPLITS * 2,), dtype=torch.float64, device=device)
    key = (device.index, stream)
    ws = _WORKSPACE.get(key)
    if ws is None or ws.numel() < c * _MAX_SPLITS * 2:
        ws = torch.empty((max(c, 256) * _MAX_SPLITS * 2,), dtype=torch.float64, device=device)
        _WORKSPACE[key] = ws
    return ws


_WORKSPACE_NHWC = {}
_NHWC_BYTES = {}
_NHWC_MAX_C = 1024


def _nhwc_workspace(device, c, stream):
    """Workspace of the channels_last kernels (per-block partial sums, backward coefficients, a ticket that must be
    zero before the first launch and is left zero by every launch)."""
    nbytes = _NHWC_BYTES.get(c)
    if nbytes is None:
        nbytes = _NHWC_BYTES[c] = _lib.bn_nhwc_workspace_bytes(c)
    if torch.cuda.is_current_stream_capturing():
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        ws[-16:].zero_()
        return ws
    key = (device.index, stream, c)
    ws = _WORKSPACE_NHWC.get(key)
    if ws is None:
        ws = torch.zeros((nbytes,), dtype=torch.uint8, device=device)
        _WORKSPACE_NHWC[key] = ws
    return ws


def _own_nhwc():
    """RVB_BN_NHWC=1: run the rvb NHWC kernels on channels_last input instead of handing it to cuDNN (1.4-1.7x faster
    kernels on the U-Net's largest layers; pays off where the host is not the bound, e.g. under CUDA-graph replay)."""
    return os.environ.get("RVB_BN_NHWC", "0") not in ("", "0")


def _is_nhwc(x):
    """4-D, dense in torch.channels_last order but not in NCHW order, with a channel count the NHWC kernels take."""
    return x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last) \
        and x.shape[1] % 4 == 0 and x.shape[1] <= _NHWC_MAX_C


class _BatchNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps):
        nhwc = _is_nhwc(x)
        if not nhwc:
            x = x.contiguous()
        n, c = x.shape[0], x.shape[1]
        hw = x.numel() // (n * c)
        y = torch.empty_like(x)                                  # keeps the memory format
        st = torch.cuda.current_stream().cuda_stream
        if training:
            if n * hw <= 1:
                raise ValueError("Expected more than 1 value per channel when training, got input size %s" % (tuple(x.shape),))
            stats = torch.empty((2, c), dtype=torch.float32, device=x.device)
            mean, invstd = stats[0], stats[1]
            if nhwc:
                _lib.call_on("rvb_bn_train_forward_nhwc", st, x.data_ptr(), n * hw, c, _ptr(weight), _ptr(bias), float(eps),
                             float(momentum), _ptr(running_mean), _ptr(running_var), mean.data_ptr(), invstd.data_ptr(),
                             y.data_ptr(), _nhwc_workspace(x.device, c, st).data_ptr())
            else:
                _lib.call_on("rvb_bn_train_forward", st, x.data_ptr(), n, c, hw, _ptr(weight), _ptr(bias), float(eps),
                             float(momentum), _ptr(running_mean), _ptr(running_var), mean.data_ptr(), invstd.data_ptr(),
                             y.data_ptr(), _partials(x.device, c, st).data_ptr())
        else:
            mean = running_mean
            invstd = torch.rsqrt(running_var + eps)
            if nhwc:
                _lib.call("rvb_bn_apply_nhwc", x.data_ptr(), n * hw, c, mean.data_ptr(), invstd.data_ptr(), _ptr(weight),
                          _ptr(bias), y.data_ptr())
            else:
                _lib.call("rvb_bn_apply", x.data_ptr(), n, c, hw, mean.data_ptr(), invstd.data_ptr(), _ptr(weight),
                          _ptr(bias), y.data_ptr())
        ctx.save_for_backward(x, weight, mean, invstd)
        ctx.training = training
        ctx.has_bias = bias is not None
        ctx.nhwc = nhwc
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, mean, invstd = ctx.saved_tensors
        dy = dy.contiguous(memory_format=torch.channels_last) if ctx.nhwc else dy.contiguous()
        n, c = x.shape[0], x.shape[1]
        hw = x.numel() // (n * c)
        need_x, need_w, need_b = ctx.needs_input_grad[0], weight is not None and ctx.needs_input_grad[1], \
            ctx.has_bias and ctx.needs_input_grad[2]
        if not (need_x or need_w or need_b):
            return (None,) * 8
        dx = torch.empty_like(x) if need_x else None
        dgb = torch.empty((2, c), dtype=torch.float32, device=x.device) if (need_w or need_b) else None
        dgamma, dbeta = (dgb[0] if need_w else None), (dgb[1] if need_b else None)
        st = torch.cuda.current_stream().cuda_stream
        if ctx.nhwc:
            _lib.call_on("rvb_bn_train_backward_nhwc", st, x.data_ptr(), dy.data_ptr(), n * hw, c, _ptr(weight),
                         mean.data_ptr(), invstd.data_ptr(), int(ctx.training), _ptr(dx), _ptr(dgamma), _ptr(dbeta),
                         _nhwc_workspace(x.device, c, st).data_ptr())
        else:
            _lib.call_on("rvb_bn_train_backward", st, x.data_ptr(), dy.data_ptr(), n, c, hw, _ptr(weight), mean.data_ptr(),
                         invstd.data_ptr(), int(ctx.training), _ptr(dx), _ptr(dgamma), _ptr(dbeta),
                         _partials(x.device, c, st).data_ptr())
        return dx, dgamma, dbeta, None, None, None, None, None


def batch_norm(x, running_mean, running_var, weight=None, bias=None, training=False, momentum=0.1, eps=1e-5):
    """``F.batch_norm`` for (N, C, ...) float32 CUDA tensors."""
    for t in (x, weight, bias, running_mean, running_var):
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise _lib.RvbError("reconvat_b200 batch_norm needs CUDA float32 tensors (got %s, %s); there is no CPU path"
                                % (t.device, t.dtype))
    if not training and (running_mean is None or running_var is None):
        raise ValueError("batch_norm in eval mode needs running_mean and running_var")
    return _BatchNormFn.apply(x, weight, bias, running_mean, running_var, bool(training), momentum, eps)


class BatchNorm2d(nn.BatchNorm2d):
    """torch.nn.BatchNorm2d (as used by model/self_attention_VAT.py:848) over the rvb_bn_* kernels."""

    def forward(self, input):
        self._check_input_dim(input)
        if not input.is_cuda or input.dtype != torch.float32:
            raise _lib.RvbError("reconvat_b200 batch_norm needs CUDA float32 tensors (got %s, %s); there is no CPU path"
                                % (input.device, input.dtype))
        if not input.is_contiguous() and _is_nhwc(input) and not _own_nhwc():
            # torch.channels_last: cuDNN has real NHWC kernels here (batchnorm_*_nhwc_semiPersist, not the
            # one-block-per-channel NCHW ones), and in eager mode ATen's per-call host cost is half that of a Python
            # autograd.Function -- the caller's iteration is host-bound in this layout (profiles/r02_train_step.md)
            return super().forward(input)
        # torch/nn/modules/batchnorm.py, _BatchNorm.forward: the exponential-average factor and the counter
        training, track = self.training, self.track_running_stats
        factor = 0.0 if self.momentum is None else self.momentum
        if training and track and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
            if self.momentum is None:
                factor = 1.0 / float(self.num_batches_tracked)
        use_running = (not training) or track
        rm = self.running_mean if use_running else None
        rv = self.running_var if use_running else None
        return _BatchNormFn.apply(input, self.weight, self.bias, rm, rv, training or (rm is None and rv is None), factor,
                                  self.eps)


def convert(module):
    """Swap the class of every ``nn.BatchNorm2d`` inside ``module`` (in place; parameters and buffers untouched)."""
    for m in module.modules():
        if type(m) is nn.BatchNorm2d:
            m.__class__ = BatchNorm2d
    return module


class _NNProxy:
    """What ``install(batchnorm=True)`` binds to the name ``nn`` inside the reference's model files: torch.nn with
    BatchNorm2d replaced."""

    BatchNorm2d = BatchNorm2d

    def __getattr__(self, name):
        return getattr(nn, name)


nn_proxy = _NNProxy()
