"""Drop-ins for the reference's VAT modules, backed by the fused row kernels of librvb.so.

The reference has the same loop five times with small differences; one core implements it and the
named classes only fix the flavour (constructor signature, call convention, scale, return arity):

==========================  ==========================================  =================================
class here                  replaces                                    differences
==========================  ==========================================  =================================
``stepwise_VAT_vatpy``      model/VAT.py:9-40                           no clamp, no scale, 2-tuple
``stepwise_VAT``            model/self_attention_VAT.py:101-145         ``model(x)``, no scale, 3-tuple
``UNet_VAT``                model/self_attention_VAT.py:147-202         ``model.transcriber(x)``, *1e10
``onset_frame_VAT``         model/self_attention_VAT.py:204-238         3-tuple model output, 2-tuple
``UNet_VAT_onset``          model/UNet_onset.py:101-162                 frame+onset heads, dict loss
``stepwise_VAT_onf``        model/onset_frame_VAT.py:158-207            3-D x, frame head = output[2]
``stepwise_VAT_frame_stack``  model/onset_frame_VAT.py:209-263          (activation, frame): MSE / BCE / both, *1e20
``Seg_VAT``                 model/Segmentation.py:22-77                 ``model(x)`` is the posterior itself, *1e10
==========================  ==========================================  =================================

``KL_Div=True`` (binary_kl_div, model/self_attention_VAT.py:248-255) and ``binwise=True`` (d / (|d| + 1e-8),
:242-243) select the matching kernels (``rvb_div_*`` with RVB_DIV_BKL, ``rvb_vat_*_binwise``).

What changes relative to the reference's op sequence (results identical within the stated tolerances):
* ``x_adv`` is made a leaf and the model's backward is driven with ``torch.autograd.grad`` seeded by
  our BCE-gradient kernel, so autograd no longer walks clamp/add/mul/div/norm and never computes
  weight gradients that ``model.zero_grad()`` would throw away (model/self_attention_VAT.py:183-185);
* the chain rule through the row normalisation and the clamp mask, the ``*1e10``, the second
  normalisation, ``eps`` scaling, the clamp of ``x + r_adv`` and the recomputed ``_l2_normalize(d)``
  are one kernel (``rvb_vat_finalize``);
* the two host-synchronising NaN asserts (:189-190) become ONE device flag written by the finalisation kernel.
  Eager calls test it synchronously, like the reference (one 4-byte read before the loss is handed back: a NaN
  ``r_adv`` never reaches ``loss.backward()`` / ``optimizer.step()``).  ``strict=False`` (or ``RVB_STRICT_NAN=0``)
  defers the test to the next call or an explicit ``check()`` -- for loops that must not synchronise every step and
  call ``vat.check()`` before ``optimizer.step()`` themselves; inside a CUDA-graph capture there is no host round
  trip at all and the owner of the graph tests ``last_flag`` (``pipeline.HotPathStep.check``).  The flavours whose
  reference has no assert (model/VAT.py, self_attention_VAT.stepwise_VAT / onset_frame_VAT) never raise.
"""
import os

import torch
import torch.nn as nn

from . import _lib

__all__ = ["stepwise_VAT_vatpy", "stepwise_VAT", "UNet_VAT", "onset_frame_VAT", "UNet_VAT_onset",
           "stepwise_VAT_onf", "stepwise_VAT_frame_stack", "Seg_VAT", "bce_mean", "binary_kl_div", "mse_mean",
           "l2_normalize", "randn_like", "Scratch"]


def _rows(x):
    return x.numel() // x.shape[-1], x.shape[-1]


def _check_input(x):
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise _lib.RvbError("reconvat_b200 VAT needs a CUDA tensor (got %s); there is no CPU path"
                            % (x.device if isinstance(x, torch.Tensor) else type(x)))
    if x.dtype != torch.float32:
        raise _lib.RvbError("reconvat_b200 VAT needs float32, got %s" % x.dtype)
    return x if x.is_contiguous() else x.contiguous()     # the reference hands over a transposed view


def _draw_direction(x_in, x):
    """``torch.randn_like(x)`` of the reference (model/self_attention_VAT.py:172) for the contiguous ``x`` the kernels
    read.  The reference draws on the tensor it was handed -- a transposed view in ``run_on_batch`` (:1104) --, and
    ``randn_like`` keeps those strides and fills in MEMORY order, so the element <-> noise mapping depends on them.
    Drawing with the incoming strides and copying afterwards reproduces the reference's ``d`` bit for bit under the
    same seed (the copy is one extra pass; contiguous inputs, what our own front-end hands over, take no copy)."""
    if x_in.is_contiguous():
        return torch.randn_like(x)
    return torch.randn_like(x_in).contiguous()


_ATEN_RANDN_LIKE = torch.randn_like      # whoever rebinds torch.randn_like (tests injecting d) keeps getting called
_max_aten_blocks = {}


def philox_geometry(numel, device):
    """ATen's launch geometry for ``normal_`` on a contiguous tensor (DistributionTemplates.h, calc_execution_policy):
    (total threads TT, what the draw adds to the generator's Philox offset)."""
    index = device.index if device.index is not None else torch.cuda.current_device()
    cap = _max_aten_blocks.get(index)
    if cap is None:
        prop = torch.cuda.get_device_properties(index)
        cap = _max_aten_blocks[index] = prop.multi_processor_count * (prop.max_threads_per_multi_processor // 256)
    tt = 256 * min(cap, -(-numel // 256))
    return tt, ((numel - 1) // (4 * tt) + 1) * 4


def _philox_args(x, rng_state):
    """(seed, offset, TT, increment, device-state pointer) for a draw of x.numel() normals.  Eager calls read the CUDA
    generator's (seed, offset) and advance it exactly as ATen would; inside a CUDA-graph capture the stream lives in
    ``rng_state`` (device uint64[3], see Scratch) and continues from replay to replay."""
    tt, inc = philox_geometry(x.numel(), x.device)
    if rng_state is not None:
        return 0, 0, tt, inc, rng_state.data_ptr()
    gen = torch.cuda.default_generators[x.device.index if x.device.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed(), gen.get_offset()
    gen.set_offset(offset + inc)
    return seed & 0xFFFFFFFFFFFFFFFF, offset, tt, inc, None


def randn_like(x, rng_state=None):
    """``torch.randn_like(x)`` for a contiguous float32 CUDA tensor, bit for bit, by ``rvb_randn_like`` (same global
    Philox stream, model/self_attention_VAT.py:172)."""
    x = _check_input(x)
    d = torch.empty_like(x)
    _lib.call("rvb_randn_like", d.data_ptr(), d.numel(), *_philox_args(x, rng_state))
    return d


def _perturb_draw(x, x_adv, d_out, n_rows, row_len, xi, clamp, rng_state=None):
    """x_adv = clamp(x + XI * d / ||d||_row) with d = torch.randn_like(x) drawn inside the row kernel, bit for bit."""
    _lib.call("rvb_vat_perturb_draw", x.data_ptr(), None if d_out is None else d_out.data_ptr(), x_adv.data_ptr(), n_rows,
              row_len, float(xi), int(clamp), *_philox_args(x, rng_state))


class _DivMean(torch.autograd.Function):
    """A divergence between posteriors with our forward and backward kernels: ``kind`` BCE
    (``F.binary_cross_entropy``), BKL (the reference's ``binary_kl_div``) or MSE (``F.mse_loss``); ``y`` is a label
    (no gradient), exactly as y_ref in the VAT loop."""

    @staticmethod
    def forward(ctx, p, y, workspace, kind, denom):
        p = p.contiguous()
        y = y.contiguous()
        loss = torch.empty((), dtype=torch.float32, device=p.device)
        _lib.call("rvb_div_mean", kind, _lib.ptr(p), _lib.ptr(y), p.numel(), float(denom), loss.data_ptr(),
                  workspace.data_ptr())
        ctx.save_for_backward(p, y)
        ctx.kind, ctx.denom = kind, denom
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        p, y = ctx.saved_tensors
        grad = torch.empty_like(p)
        go = grad_out.contiguous().to(torch.float32)
        _lib.call("rvb_div_grad", ctx.kind, p.data_ptr(), y.data_ptr(), grad.data_ptr(), p.numel(), float(ctx.denom),
                  go.data_ptr(), 1.0)
        return grad, None, None, None, None


_workspaces = {}


def _workspace(device):
    """Default reduction workspace (ticket + per-block partials) of the divergence kernels: one per (device, current
    stream).  Calls on one stream are ordered, so they may share it; calls in flight on different streams (side
    streams, DataParallel threads) must not -- they would race on the ticket.  A zero-filled workspace is allocated
    on the calling stream, so its memset is ordered before the first kernel that uses it."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.RvbError("reconvat_b200: the first divergence call on a stream allocates its reduction workspace; "
                                "run the step once eagerly on this stream (or attach a VAT.Scratch) before capturing")
        ws = torch.zeros(_lib.BCE_WORKSPACE_FLOATS, dtype=torch.float32, device=device)
        _workspaces[key] = ws
    return ws


def _denom(p, kind):
    """What the reference divides the summed divergence by: numel for the means, the batch size for
    ``F.kl_div(..., reduction='batchmean')``."""
    return p.shape[0] if kind == _lib.DIV_BKL else p.numel()


def _divergence(p, y, kind, workspace=None):
    if p.shape != y.shape:
        raise ValueError("Using a target size ({}) that is different to the input size ({}) is deprecated. "
                         "Please ensure they have the same size.".format(y.shape, p.shape))
    ws = _workspace(p.device) if workspace is None else workspace
    return _DivMean.apply(p, y.detach(), ws, kind, _denom(p, kind))


class Scratch:
    """Reduction workspaces of ONE in-flight VAT call chain.  The divergence and the finalisation kernels finish with
    a last-block reduction over a small workspace (ticket + per-block partials); two calls that may run concurrently
    -- CUDA graphs replayed on different streams -- must not share it.  ``vat.scratch = Scratch(device)`` also
    switches the module to ``rvb_vat_finalize_stats``: the NaN flag is written (not OR-ed into a zeroed one) and
    ``vat.last_r_norm_mean`` receives mean |d_hat| (the ``r_norm.abs().mean()`` of model/self_attention_VAT.py:1149)
    without another pass over d_hat."""

    def __init__(self, device, keep_d_hat=True):
        self.device = torch.device(device)
        self.div = torch.zeros(_lib.BCE_WORKSPACE_FLOATS, dtype=torch.float32, device=self.device)
        self._stats = None
        # a private Philox stream for the in-kernel draw of d inside captured graphs: {seed, offset, ticket}.  It starts
        # at the CUDA generator's current position, and the generator skips 2^40 draws so that neither it nor another
        # Scratch ever revisits this segment (replays advance the device-side offset, never the host generator).
        gen = torch.cuda.default_generators[self.device.index if self.device.index is not None
                                            else torch.cuda.current_device()]
        seed, offset = gen.initial_seed(), gen.get_offset()
        gen.set_offset(offset + (1 << 40))
        self.rng_state = torch.tensor([seed - (1 << 64) if seed >= (1 << 63) else seed, offset, 0], dtype=torch.int64,
                                      device=self.device)
        # False: the module returns None for its third output (the normalised direction) and the kernel does not
        # store it -- for callers that only log its mean (``last_r_norm_mean``), as run_on_batch does
        self.keep_d_hat = keep_d_hat

    def stats(self, n_rows):
        need = _lib.vat_stats_workspace_bytes(n_rows)
        if self._stats is None or self._stats.numel() * 4 < need:
            if torch.cuda.is_current_stream_capturing():
                raise _lib.RvbError("reconvat_b200 VAT: the Scratch of a captured step must be sized before the capture "
                                    "(run the step once eagerly with the same scratch)")
            self._stats = torch.zeros((need + 3) // 4, dtype=torch.int32, device=self.device)
        return self._stats


def bce_mean(p, y):
    """Differentiable (w.r.t. ``p``) mean binary cross entropy on the device kernels."""
    return _divergence(p, y, _lib.DIV_BCE)


def binary_kl_div(y_pred, y_ref):
    """model/self_attention_VAT.py:248-255 on the device kernels (differentiable w.r.t. ``y_pred``)."""
    return _divergence(y_pred, y_ref, _lib.DIV_BKL)


def mse_mean(p, y):
    """``F.mse_loss(p, y)`` on the device kernels (differentiable w.r.t. ``p``)."""
    return _divergence(p, y, _lib.DIV_MSE)


def _div_grad(p, y, kind):
    """d divergence / d p as a plain tensor (seeds the model's backward in the power iteration)."""
    p = p.detach().contiguous()
    y = y.detach().contiguous()
    grad = torch.empty_like(p)
    _lib.call("rvb_div_grad", kind, _lib.ptr(p), _lib.ptr(y), grad.data_ptr(), p.numel(), float(_denom(p, kind)),
              None, 1.0)
    return grad


def l2_normalize(d):
    """``_l2_normalize(d, binwise=False)`` (model/self_attention_VAT.py:240-246) as one kernel."""
    d = _check_input(d)
    n_rows, row_len = _rows(d)
    out = torch.empty_like(d)
    scratch = torch.empty_like(d)
    _lib.call("rvb_vat_direct", d.data_ptr(), d.data_ptr(), scratch.data_ptr(), scratch.data_ptr(), out.data_ptr(),
              n_rows, row_len, 1.0, 0, None)
    return out


class _VATCore(nn.Module):
    # flavour knobs, overridden by the named subclasses
    _use_transcriber = False       # model.transcriber(x) vs model(x)
    _heads = (0,)                  # indices of the model outputs that enter the divergence (None: the output itself)
    _kinds = None                  # divergence per head; None -> BCE, or binary KL when KL_Div=True
    _scale = 1.0                   # d.grad multiplier (1e10 in the UNet / O&F flavours)
    _clamp = True                  # (x + r).clamp(0, 1)
    _n_returns = 3
    _dict_loss = None              # names of the per-head losses when the loss is returned as a dict
    _asserts = True                # the reference flavour asserts on NaN / Inf in r_adv (False: it never looks)
    _nan_message = ("r_adv has nan, d min={dmin} d max={dmax} d mean={dmean} please debug tune down the XI for VAT")

    def _init_common(self, XI, epsilon, n_power, KL_Div=False, binwise=False, strict=None):
        self.n_power = n_power
        self.XI = XI
        self.epsilon = epsilon
        self.KL_Div = KL_Div
        self.binwise = binwise
        self.strict = bool(int(os.environ.get("RVB_STRICT_NAN", "1"))) if strict is None else bool(strict)
        self._pending = None       # (pinned host copy of the NaN flag, event) of the previous eager call
        self._host_flag = None     # persistent pinned int32 (allocated once: no per-call cudaHostAlloc)
        self.last_flag = None      # device flag of the latest call (what a captured CUDA graph leaves behind)
        self.scratch = None        # a Scratch: private reduction workspaces + fused flag / mean |d_hat| (see Scratch)
        self.last_r_norm_mean = None   # device scalar, mean |d_hat| of the latest call (only with a Scratch)
        if KL_Div and len(self._heads) > 1:
            raise NotImplementedError("reconvat_b200 VAT: KL_Div=True with two heads -- the reference itself fails "
                                      "there (NameError: y_pred, model/UNet_onset.py:133-134)")
        if n_power not in (0, 1):
            raise NotImplementedError("reconvat_b200 VAT: n_power=%r -- the reference itself fails for n_power > 1 "
                                      "(d.grad is None on the second iteration, model/self_attention_VAT.py:184)"
                                      % (n_power,))

    # -- deferred NaN / Inf assertion ---------------------------------------------------------
    def check(self, flag=None):
        """Raise the reference's AssertionError if the previous call produced NaN/Inf in r_adv.
        ``flag``: a device flag to test instead (e.g. ``last_flag`` after replaying a captured graph; this
        synchronises with the device)."""
        if not self._asserts:
            self._pending = None
            return
        if flag is not None:
            bad = int(flag.item()) != 0
        elif self._pending is not None:
            host, event = self._pending
            self._pending = None
            event.synchronize()
            bad = int(host.item()) != 0
        else:
            return
        if bad:
            raise AssertionError(self._nan_message.format(dmin="nan", dmax="nan", dmean="nan"))

    def _model_outputs(self, model, x):
        out = model.transcriber(x) if self._use_transcriber else model(x)
        return [out if i is None else out[i] for i in self._heads]

    def _head_kinds(self):
        if self._kinds is not None:
            return self._kinds
        return (_lib.DIV_BKL if self.KL_Div else _lib.DIV_BCE,) * len(self._heads)

    def forward(self, model, x):
        x_in = x
        x = _check_input(x).detach()
        if not torch.cuda.is_current_stream_capturing():
            self.check()
        n_rows, row_len = _rows(x)
        with torch.no_grad():
            y_ref = [y.detach() for y in self._model_outputs(model, x)]   # labels, no grad (…:163-164)

        capturing = torch.cuda.is_current_stream_capturing()
        # d ~ N(0, 1) from the global Philox stream (…:172): drawn inside the perturb kernel when x is contiguous
        # (bit-identical to torch.randn_like), by ATen on the caller's strides otherwise
        # RVB_DRAW = kernel (default): rvb_randn_like, ATen's amortisation; fused: inside rvb_vat_perturb (no round trip
        # of d, 4x the Philox work: fastest on an otherwise idle GPU); aten: torch.randn_like itself
        mode = os.environ.get("RVB_DRAW", "kernel")
        ours = (mode != "aten" and x_in.is_contiguous() and torch.randn_like is _ATEN_RANDN_LIKE
                and (not capturing or self.scratch is not None))
        fused_draw = ours and mode == "fused" and self.n_power == 1 and not self.binwise
        rng_state = self.scratch.rng_state if (capturing and self.scratch is not None) else None
        d = torch.empty_like(x) if fused_draw else (randn_like(x, rng_state) if ours else _draw_direction(x_in, x))
        sc = self.scratch if not self.binwise else None
        div_ws = None if self.scratch is None else self.scratch.div
        if sc is not None:
            flag = torch.empty((), dtype=torch.int32, device=x.device)    # written by the kernel's last block
            r_norm_mean = torch.empty((), dtype=torch.float32, device=x.device)
            stats_ws = sc.stats(n_rows)
        else:
            flag = torch.zeros((), dtype=torch.int32, device=x.device)
        r_adv = torch.empty_like(x)
        x_adv2 = torch.empty_like(x)
        d_hat = torch.empty_like(x) if (sc is None or sc.keep_d_hat) else None
        d_hat_ptr = None if d_hat is None else d_hat.data_ptr()
        kinds = self._head_kinds()
        if self.n_power == 1:
            x_adv = torch.empty_like(x)
            if self.binwise:
                _lib.call("rvb_vat_perturb_binwise", x.data_ptr(), d.data_ptr(), x_adv.data_ptr(), x.numel(),
                          float(self.XI), int(self._clamp))
            elif fused_draw:
                _perturb_draw(x, x_adv, d, n_rows, row_len, self.XI, self._clamp, rng_state)
            else:
                _lib.call("rvb_vat_perturb", x.data_ptr(), d.data_ptr(), x_adv.data_ptr(), n_rows, row_len,
                          float(self.XI), int(self._clamp))
            x_adv.requires_grad_(True)
            with torch.enable_grad():
                y_pred = self._model_outputs(model, x_adv)
                seeds = [_div_grad(p, y, k) for p, y, k in zip(y_pred, y_ref, kinds)]  # d(sum of divergences)/dp (…:182)
                (g,) = torch.autograd.grad(y_pred, [x_adv], seeds)        # model backward only (…:183)
            model.zero_grad()                                             # side effect kept (…:185)
            if self.binwise:
                _lib.call("rvb_vat_finalize_binwise", g.contiguous().data_ptr(), d.data_ptr(), x.data_ptr(),
                          r_adv.data_ptr(), x_adv2.data_ptr(), d_hat.data_ptr(), x.numel(), float(self.XI),
                          float(self.epsilon), float(self._scale), int(self._clamp), flag.data_ptr())
            elif sc is not None:
                _lib.call("rvb_vat_finalize_stats", g.contiguous().data_ptr(), d.data_ptr(), x.data_ptr(),
                          r_adv.data_ptr(), x_adv2.data_ptr(), d_hat_ptr, n_rows, row_len, float(self.XI),
                          float(self.epsilon), float(self._scale), int(self._clamp), flag.data_ptr(),
                          r_norm_mean.data_ptr(), stats_ws.data_ptr(), stats_ws.numel() * 4)
            else:
                _lib.call("rvb_vat_finalize", g.contiguous().data_ptr(), d.data_ptr(), x.data_ptr(), r_adv.data_ptr(),
                          x_adv2.data_ptr(), d_hat.data_ptr(), n_rows, row_len, float(self.XI), float(self.epsilon),
                          float(self._scale), int(self._clamp), flag.data_ptr())
        elif self.binwise:
            _lib.call("rvb_vat_finalize_binwise", None, d.data_ptr(), x.data_ptr(), r_adv.data_ptr(), x_adv2.data_ptr(),
                      d_hat.data_ptr(), x.numel(), float(self.XI), float(self.epsilon), 1.0, int(self._clamp),
                      flag.data_ptr())
        elif sc is not None:
            _lib.call("rvb_vat_finalize_stats", None, d.data_ptr(), x.data_ptr(), r_adv.data_ptr(), x_adv2.data_ptr(),
                      d_hat_ptr, n_rows, row_len, 0.0, float(self.epsilon), 1.0, int(self._clamp),
                      flag.data_ptr(), r_norm_mean.data_ptr(), stats_ws.data_ptr(), stats_ws.numel() * 4)
        else:
            _lib.call("rvb_vat_direct", d.data_ptr(), x.data_ptr(), r_adv.data_ptr(), x_adv2.data_ptr(),
                      d_hat.data_ptr(), n_rows, row_len, float(self.epsilon), int(self._clamp), flag.data_ptr())

        self.last_flag = flag
        self.last_r_norm_mean = r_norm_mean if sc is not None else None
        if self._asserts and not torch.cuda.is_current_stream_capturing():
            # inside a CUDA-graph capture there is no host round trip: the owner of the graph tests last_flag
            if self._host_flag is None:
                self._host_flag = torch.empty((), dtype=torch.int32, pin_memory=True)
            self._host_flag.copy_(flag, non_blocking=True)
            event = torch.cuda.Event()
            event.record()
            self._pending = (self._host_flag, event)
            if self.strict:
                self.check()

        y_pred = self._model_outputs(model, x_adv2)                       # graph to the parameters kept (…:195)
        losses = [_divergence(p, y, k, div_ws) for p, y, k in zip(y_pred, y_ref, kinds)]   # (…:200)
        if self._dict_loss is not None:
            vat_loss = dict(zip(self._dict_loss, losses))
        else:
            vat_loss = losses[0]
            for extra in losses[1:]:
                vat_loss = vat_loss + extra
        if self._n_returns == 2:
            return vat_loss, r_adv
        return vat_loss, r_adv, d_hat


class stepwise_VAT_vatpy(_VATCore):
    """model/VAT.py:9-40 -- ``stepwise_VAT(XI, epsilon, n_power)``; no clamp, no assert, returns (vat_loss, r_adv)."""
    _clamp = False
    _n_returns = 2
    _asserts = False

    def __init__(self, XI, epsilon, n_power, strict=None):
        super().__init__()
        self._init_common(XI, epsilon, n_power, strict=strict)


class stepwise_VAT(_VATCore):
    """model/self_attention_VAT.py:101-145 -- ``stepwise_VAT(XI, epsilon, n_power, KL_Div, binwise=False)``; the
    reference has no NaN assert in this flavour."""
    _asserts = False

    def __init__(self, XI, epsilon, n_power, KL_Div, binwise=False, strict=None):
        super().__init__()
        self._init_common(XI, epsilon, n_power, KL_Div, binwise, strict)


class UNet_VAT(_VATCore):
    """model/self_attention_VAT.py:147-202 -- ``UNet_VAT(XI, epsilon, n_power, KL_Div, reconstruction=False)``."""
    _use_transcriber = True
    _scale = 1e10

    def __init__(self, XI, epsilon, n_power, KL_Div, reconstruction=False, strict=None):
        super().__init__()
        self._init_common(XI, epsilon, n_power, KL_Div, False, strict)
        self.reconstruction = reconstruction


class onset_frame_VAT(_VATCore):
    """model/self_attention_VAT.py:204-238 -- ``onset_frame_VAT(XI, epsilon, n_power)``; the model returns a
    3-tuple whose first element is the posterior; returns (vat_loss, r_adv); no NaN assert in the reference."""
    _n_returns = 2
    _asserts = False

    def __init__(self, XI, epsilon, n_power, strict=None):
        super().__init__()
        self._init_common(XI, epsilon, n_power, strict=strict)


class UNet_VAT_onset(_VATCore):
    """model/UNet_onset.py:101-162 -- frame and onset posteriors, loss returned as {'frame','onset'}."""
    _use_transcriber = True
    _heads = (0, 1)
    _scale = 1e10
    _dict_loss = ("frame", "onset")

    def __init__(self, XI, epsilon, n_power, KL_Div, reconstruction=False, strict=None):
        super().__init__()
        self._init_common(XI, epsilon, n_power, KL_Div, False, strict)
        self.reconstruction = reconstruction


class stepwise_VAT_onf(_VATCore):
    """model/onset_frame_VAT.py:158-207 -- ``stepwise_VAT(XI, epsilon, n_power, KL_Div)`` of the Onsets&Frames
    baseline: x is (B, T, F), the model returns (onset, activation, frame) and only ``frame`` enters the loss;
    the third return is ``_l2_normalize(d*1e8)`` == ``_l2_normalize(d)`` away from overflow."""
    _heads = (2,)
    _scale = 1e10
    _nan_message = "r_adv contains nan"

    def __init__(self, XI, epsilon, n_power, KL_Div, strict=None):
        super().__init__()
        self._init_common(XI, epsilon, n_power, KL_Div, False, strict)


class stepwise_VAT_frame_stack(_VATCore):
    """model/onset_frame_VAT.py:209-263 -- ``stepwise_VAT_frame_stack(XI, epsilon, n_power, VAT_mode)``: the model
    returns (activation, frame); the divergence is MSE on the activation ('activation'), BCE on the frame posterior
    ('frame') or their sum ('all'); d.grad is scaled by 1e20; returns (vat_loss, r_adv)."""
    _scale = 1e20
    _n_returns = 2
    _nan_message = "r_adv exploded, please debug tune down the XI for VAT"

    def __init__(self, XI, epsilon, n_power, VAT_mode, strict=None):
        super().__init__()
        modes = {"activation": ((0,), (_lib.DIV_MSE,)), "frame": ((1,), (_lib.DIV_BCE,)),
                 "all": ((1, 0), (_lib.DIV_BCE, _lib.DIV_MSE))}          # dist_frame + dist_activation (:241)
        if VAT_mode not in modes:
            # the reference leaves `dist` unbound for any other mode (UnboundLocalError at :243)
            raise ValueError("VAT_mode must be 'activation', 'frame' or 'all', got %r" % (VAT_mode,))
        self._heads, self._kinds = modes[VAT_mode]
        self.VAT_mode = VAT_mode
        self._init_common(XI, epsilon, n_power, strict=strict)


class Seg_VAT(_VATCore):
    """model/Segmentation.py:22-77 -- ``Seg_VAT(XI, epsilon, n_power, KL_Div, reconstruction=False)``: ``model(x)`` is
    the posterior itself (no tuple), d.grad * 1e10, returns (vat_loss, r_adv, d_hat)."""
    _heads = (None,)
    _scale = 1e10

    def __init__(self, XI, epsilon, n_power, KL_Div, reconstruction=False, strict=None):
        super().__init__()
        if KL_Div:
            raise NotImplementedError("reconvat_b200 Seg_VAT: KL_Div=True -- the reference itself fails there "
                                      "(NameError: binary_kl_div is not defined in model/Segmentation.py:55)")
        self._init_common(XI, epsilon, n_power, KL_Div, False, strict)
        self.reconstruction = reconstruction
