"""Drop-in for ``model.utils.Normalization`` (model/utils.py:82-106), 'imagewise' mode on the device
kernels: per-sample min and max over all (bins x frames) values, then (x - min) / (max - min).
No epsilon, like the reference: a constant image becomes NaN."""
import torch

from . import _lib


class Normalization():
    def __init__(self, mode='framewise'):
        if mode == 'imagewise':
            self.normalize = _imagewise
        elif mode == 'framewise':
            # model/utils.py:85-92; not selected by any shipped script (all pass mode='imagewise',
            # train_UNet_VAT.py:19) -> outside the accelerated path, refuse rather than emulate.
            def normalize(x):
                raise NotImplementedError("reconvat_b200.Normalization: only mode='imagewise' is accelerated")
            self.normalize = normalize
        else:
            print(f'please choose the correct mode')
        self.mode = mode

    def transform(self, x):
        return self.normalize(x)


def _imagewise(x):
    if not x.is_cuda or x.dtype != torch.float32:
        raise _lib.RvbError("reconvat_b200.Normalization needs a CUDA float32 tensor (got %s, %s); "
                            "there is no CPU path" % (x.device, x.dtype))
    if x.dim() != 3:
        raise ValueError("Normalization('imagewise') expects (batch, bins, frames)")
    xc = x.contiguous()
    B = xc.shape[0]
    n = xc.shape[1] * xc.shape[2]
    minmax = torch.empty((B, 2), dtype=torch.int32, device=x.device)
    out = torch.empty_like(xc)
    _lib.call("rvb_minmax", xc.data_ptr(), B, n, minmax.data_ptr())
    _lib.call("rvb_normalise", xc.data_ptr(), out.data_ptr(), B, n, minmax.data_ptr())
    return out
