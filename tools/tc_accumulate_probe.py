"""How does tcgen05.mma add into its fp32 accumulator?  (evidence for DESIGN.md section 2, "precision")

Runs exact little problems through ``rvb_gemm_nt_tf32x3`` (kind::tf32, K = 8 per MMA, lo planes zero so that only the
hi * hi MMAs contribute) whose true results sit BETWEEN fp32 grid points, and prints what comes back:

  cross   acc = V from one MMA, a later MMA adds s             -> rounding of the accumulate step
  intra   V and s inside ONE MMA (same 8-term block)           -> is the sum inside an MMA exact before it is rounded?
  intra7  V and seven terms s inside one MMA
  chain   acc = V, then 63 MMAs add s each                     -> drift of a long chain (round-to-nearest: none)

V = +1 or -1, s = m * 2^-26 (the fp32 grid is 8 * 2^-26 above 1, 4 * 2^-26 below).  Output: (result - V) / 2^-26 per m,
next to the exact value and to what round-to-nearest / toward-zero / toward -inf would give for the cross case.
"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from reconvat_b200 import _lib  # noqa: E402

U = 2.0 ** -26


K = 512
N_TERMS = {"cross": 1, "intra": 1, "intra7": 7, "chain": 63}


def problem():
    """(a (rows, K), b (len(ms), K), rows [(kind, V)], ms): out[i][j] = sum_k a[i][k] b[j][k]."""
    ms = np.arange(-12, 13)
    kinds = ["cross", "intra", "intra7", "chain"]
    rows = [(kind, v) for kind in kinds for v in (1.0, -1.0)]
    a = np.zeros((len(rows), K), np.float32)
    for i, (kind, v) in enumerate(rows):
        a[i, 0] = v
        if kind == "cross":
            a[i, 8] = 1.0
        elif kind == "intra":
            a[i, 1] = 1.0
        elif kind == "intra7":
            a[i, 1:8] = 1.0
        else:
            a[i, 8::8] = 1.0                      # one term in each of the 63 later MMAs
    b = np.zeros((len(ms), K), np.float32)
    b[:, 0] = 1.0
    b[:, 1:] = (ms * U)[:, None]
    return a, b, rows, ms


def run(dev):
    """The probe through rvb_gemm_nt_tf32x3 on `dev`: float32 (rows, len(ms))."""
    a, b, rows, ms = problem()
    ah = torch.from_numpy(a).to(dev); bh = torch.from_numpy(b).to(dev)
    out = torch.empty((len(rows), len(ms)), dtype=torch.float32, device=dev)
    al, bl = torch.zeros_like(ah), torch.zeros_like(bh)
    _lib.call("rvb_gemm_nt_tf32x3", ah.data_ptr(), al.data_ptr(), len(rows), bh.data_ptr(), bl.data_ptr(), len(ms), K,
              out.data_ptr(), out.stride(0), 1, 0)           # k_split = 1: ONE accumulator per element, 64 k-blocks of 8
    torch.cuda.synchronize()
    return out.cpu().numpy()


def main():
    a, b, rows, ms = problem()
    n_terms = N_TERMS
    res = run(torch.device("cuda:0")).astype(np.float64)
    if len(sys.argv) > 1:                                      # golden vector for the CPU model test
        import json
        with open(sys.argv[1], "w") as f:
            json.dump({"rows": [[k, v] for k, v in rows], "m": ms.tolist(),
                       "result_minus_v_in_units_of_2^-26": ((res - np.array([v for _, v in rows])[:, None]) / U).tolist()}, f)

    def grid(x, mode):
        f = np.float32(x)
        if mode == "rn" or float(f) == x:
            return float(f)
        lo, hi = (f, np.nextafter(f, np.float32(np.inf))) if float(f) < x else (np.nextafter(f, np.float32(-np.inf)), f)
        return float(lo if mode == "floor" else (lo if x > 0 else hi))

    print("m (s = m * 2^-26):      " + " ".join("%4d" % m for m in ms))
    for i, (kind, v) in enumerate(rows):
        n = n_terms[kind]
        print("%-6s V=%+d  result   : " % (kind, v) + " ".join("%4g" % ((res[i, j] - v) / U) for j in range(len(ms))))
        print("               exact    : " + " ".join("%4g" % (n * m) for m in ms))
        if kind == "cross":
            for mode in ("rn", "rz", "floor"):
                print("               %-9s: " % mode + " ".join("%4g" % ((grid(v + m * U, mode) - v) / U) for m in ms))
        if kind == "chain":
            def chain_of(mode):
                vals = []
                for m in ms:
                    acc = v
                    for _ in range(n):
                        acc = grid(acc + m * U, mode)
                    vals.append((acc - v) / U)
                return vals
            for mode in ("rn", "rz", "floor"):
                print("               %-9s: " % mode + " ".join("%4g" % x for x in chain_of(mode)))


if __name__ == "__main__":
    main()
