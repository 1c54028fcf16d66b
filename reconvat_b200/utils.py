"""Drop-in for ``model.utils.Normalization`` (model/utils.py:82-106) on the device kernels.  'imagewise' (what every
shipped script selects): per-sample min and max over all (bins x frames) values, then (x - min) / (max - min); no
epsilon, like the reference: a constant image becomes NaN.  'framewise': the same per (sample, frame) over the bins,
with NaN replaced by 0."""
import torch

from . import _lib


class Normalization():
    def __init__(self, mode='framewise'):
        if mode == 'imagewise':
            self.normalize = _imagewise
        elif mode == 'framewise':
            self.normalize = _framewise
        else:
            print(f'please choose the correct mode')
        self.mode = mode

    def transform(self, x):
        return self.normalize(x)


def _check(x, mode):
    if not x.is_cuda or x.dtype != torch.float32:
        raise _lib.RvbError("reconvat_b200.Normalization needs a CUDA float32 tensor (got %s, %s); "
                            "there is no CPU path" % (x.device, x.dtype))
    if x.dim() != 3:
        raise ValueError("Normalization(%r) expects (batch, bins, frames)" % mode)
    return x.contiguous()


def _framewise(x):
    xc = _check(x, 'framewise')
    out = torch.empty_like(xc)
    _lib.call("rvb_normalise_framewise", xc.data_ptr(), out.data_ptr(), xc.shape[0], xc.shape[1], xc.shape[2])
    return out


def _imagewise(x):
    xc = _check(x, 'imagewise')
    B = xc.shape[0]
    n = xc.shape[1] * xc.shape[2]
    minmax = torch.empty((B, 2), dtype=torch.int32, device=x.device)
    out = torch.empty_like(xc)
    _lib.call("rvb_minmax", xc.data_ptr(), B, n, minmax.data_ptr())
    _lib.call("rvb_normalise", xc.data_ptr(), out.data_ptr(), B, n, minmax.data_ptr())
    return out
