"""BatchNorm2d forward + backward (training) on the U-Net's own tensor shapes: reconvat_b200.batchnorm against
torch.nn.BatchNorm2d (cuDNN) on the same GPU, with the achieved HBM bandwidth of ours.  Algorithmic bytes per element:
forward 3 x 4 (read x twice, write y), backward 5 x 4 (read x and dy twice, write dx)."""
import copy
import json
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from reconvat_b200 import batchnorm  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    fmt = torch.channels_last if len(sys.argv) > 2 and sys.argv[2] == "channels_last" else torch.contiguous_format
    os.environ["RVB_BN_NHWC"] = "1"                              # measure OUR channels_last kernels, not the delegation
    peak = None
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print("B = %d; HBM peak %s GB/s; memory format %s" % (B, peak, fmt))
    print("%-22s %10s %10s %8s %10s %10s %8s %12s" % ("shape", "fwd ours", "fwd torch", "x", "bwd ours", "bwd torch", "x", "ours GB/s f/b"))
    # several tensors per shape so that consecutive calls do not find their input in the L2
    for c, h, w in ((16, 640, 229), (32, 320, 114), (64, 160, 57), (128, 80, 28), (96, 160, 57), (48, 320, 114)):
        ref = nn.BatchNorm2d(c).to(dev)
        ours = batchnorm.convert(copy.deepcopy(ref))
        n_rot = max(2, int(300e6 // (B * c * h * w * 4)) + 1)
        xs = [torch.randn(B, c, h, w, device=dev).contiguous(memory_format=fmt) for _ in range(n_rot)]
        dys = [torch.randn(B, c, h, w, device=dev).contiguous(memory_format=fmt) for _ in range(n_rot)]
        res = {}
        for name, m in (("torch", ref), ("ours", ours)):
            it = [0]

            def fwd():
                it[0] += 1
                return m(xs[it[0] % n_rot])
            t_f = timed(fwd)
            ys = []
            for i in range(n_rot):
                xi = xs[i].clone(memory_format=torch.preserve_format).requires_grad_(True)
                ys.append((m(xi), xi))

            def bwd():
                it[0] += 1
                y, xi = ys[it[0] % n_rot]
                torch.autograd.grad(y, (xi, m.weight, m.bias), dys[it[0] % n_rot], retain_graph=True)
            t_b = timed(bwd)
            res[name] = (t_f, t_b)
            del ys
        el = B * c * h * w * 4
        print("%-22s %8.1f us %8.1f us %7.1fx %8.1f us %8.1f us %7.1fx %6.0f / %-6.0f" % (
            "(%d, %d, %d, %d)" % (B, c, h, w), res["ours"][0] * 1e3, res["torch"][0] * 1e3, res["torch"][0] / res["ours"][0],
            res["ours"][1] * 1e3, res["torch"][1] * 1e3, res["torch"][1] / res["ours"][1],
            3 * el / res["ours"][0] / 1e6, 5 * el / res["ours"][1] / 1e6))


if __name__ == "__main__":
    main()
