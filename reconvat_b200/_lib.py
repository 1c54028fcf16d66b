"""ctypes binding of librvb.so (the C ABI declared in include/rvb.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``make -C
reconvat_b200/csrc``).  There is NO fallback: if the shared object is missing,
or a tensor is not a CUDA float32 tensor, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librvb.so")

_c_p = ctypes.c_void_p          # device pointers and the stream travel as integers
_i32, _i64, _f32 = ctypes.c_int, ctypes.c_int64, ctypes.c_float

# name -> argtypes; mirrors include/rvb.h one to one (tests/test_abi.py checks the header against this)
SIGNATURES = {
    "rvb_pad_split": [_c_p, _i64, _i32, _i32, _i32, _i32, _c_p, _c_p, _i32, _i32, _c_p],
    "rvb_stft_gemm": [_c_p, _c_p, _i32, _i32, _i32, _i32, _c_p, _c_p, _i32, _i32, _i32, _f32, _c_p, _i32, _c_p],
    "rvb_fold_split": [_c_p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _c_p, _c_p, _c_p, _c_p],
    "rvb_stft_gemm_folded": [_c_p, _c_p, _i32, _i32, _i32, _c_p, _c_p, _i32, _c_p, _f32, _i32, _f32, _c_p, _i32, _c_p],
    "rvb_stft_bin_folded": [_c_p, _c_p, _i32, _i32, _i32, _c_p, _c_p, _c_p, _f32, _i32, _i32, _f32, _c_p, _i32, _c_p],
    "rvb_fold_split_f16": [_c_p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _c_p, _c_p, _c_p, _c_p, _c_p],
    "rvb_fold_split_f16_pcm16": [_c_p, _i64, _f32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _c_p, _c_p, _c_p, _c_p, _c_p],
    "rvb_stft_gemm_folded_f16": [_c_p, _c_p, _c_p, _i32, _i32, _i32, _c_p, _c_p, _f32, _i32, _c_p, _f32, _i32, _f32, _c_p,
                                 _i32, _c_p],
    "rvb_stft_mel_folded_f16": [_c_p, _c_p, _c_p, _i32, _i32, _i32, _c_p, _c_p, _f32, _i32, _c_p, _f32, _i32, _f32, _c_p,
                                _i32, _c_p, _c_p],
    "rvb_fold_split2_f16": [_c_p, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _c_p, _c_p, _c_p, _c_p],
    "rvb_fold_split2_f16_pcm16": [_c_p, _i64, _f32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _c_p, _c_p, _c_p, _c_p],
    "rvb_stft_mel_folded2_f16": [_c_p, _c_p, _c_p, _i32, _i32, _i32, _c_p, _c_p, _f32, _c_p, _i32, _c_p, _c_p],
    "rvb_pad_parity_pcm16": [_c_p, _i64, _i32, _i32, _i32, _i32, _c_p, _i64, _c_p],
    "rvb_stft_mel_fused_pcm16": [_c_p, _i64, _i32, _i32, _i32, _i32, _f32, _c_p, _c_p, _f32, _c_p, _i32, _c_p, _c_p],
    "rvb_logmel_minmax": [_c_p, _i32, _i64, _f32, _c_p, _c_p],
    "rvb_logmel_transpose": [_c_p, _i32, _i32, _i32, _f32, _c_p, _c_p, _c_p],
    "rvb_logmel_normalise": [_c_p, _c_p, _i32, _i32, _i32, _f32, _c_p, _c_p, _c_p],
    "rvb_stft_bin_folded_f16": [_c_p, _c_p, _c_p, _i32, _i32, _i32, _c_p, _c_p, _c_p, _f32, _i32, _i32, _f32, _c_p, _i32,
                                _c_p],
    "rvb_stft_bin": [_c_p, _c_p, _i32, _i32, _i32, _i32, _c_p, _c_p, _i32, _i32, _i32, _f32, _c_p, _i32, _c_p],
    "rvb_mel_project": [_c_p, _i32, _i32, _i32, _c_p, _c_p, _c_p, _i32, _i32, _f32, _i32, _c_p, _c_p, _c_p],
    "rvb_minmax": [_c_p, _i32, _i64, _c_p, _c_p],
    "rvb_normalise": [_c_p, _c_p, _i32, _i64, _c_p, _c_p],
    "rvb_normalise_framewise": [_c_p, _c_p, _i32, _i32, _i32, _c_p],
    "rvb_vat_perturb": [_c_p, _c_p, _c_p, _i64, _i32, _f32, _i32, _c_p],
    "rvb_vat_perturb_draw": [_c_p, _c_p, _c_p, _i64, _i32, _f32, _i32, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                             ctypes.c_uint64, _c_p, _c_p],
    "rvb_randn_like": [_c_p, _i64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, _c_p, _c_p],
    "rvb_bce_grad": [_c_p, _c_p, _c_p, _i64, _c_p, _f32, _c_p],
    "rvb_vat_finalize": [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _i64, _i32, _f32, _f32, _f32, _i32, _c_p, _c_p],
    "rvb_div_grad": [_i32, _c_p, _c_p, _c_p, _i64, ctypes.c_double, _c_p, _f32, _c_p],
    "rvb_div_mean": [_i32, _c_p, _c_p, _i64, ctypes.c_double, _c_p, _c_p, _c_p],
    "rvb_vat_perturb_binwise": [_c_p, _c_p, _c_p, _i64, _f32, _i32, _c_p],
    "rvb_vat_finalize_binwise": [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _i64, _f32, _f32, _f32, _i32, _c_p, _c_p],
    "rvb_split_tf32": [_c_p, _i64, _i32, _i64, _i32, _c_p, _c_p, _i64, _i64, _i64, _c_p],
    "rvb_gemm_nt_tf32x3": [_c_p, _c_p, _i64, _c_p, _c_p, _i32, _i32, _c_p, _i64, _i32, _i64, _c_p],
    "rvb_local_attn_fwd": [_c_p, _c_p, _c_p, _c_p, _i32, _i32, _i32, _i32, _i32, _c_p, _c_p, _c_p],
    "rvb_local_attn_bwd_q": [_c_p, _c_p, _c_p, _c_p, _i32, _i32, _i32, _i32, _i32, _c_p, _c_p, _c_p],
    "rvb_local_attn_bwd_kv": [_c_p, _c_p, _c_p, _c_p, _i32, _i32, _i32, _i32, _i32, _c_p, _c_p, _c_p],
    "rvb_note_offsets": [_c_p, _c_p, _i32, _i32, _f32, _f32, _i32, _c_p, _c_p, _c_p],
    "rvb_vat_direct": [_c_p, _c_p, _c_p, _c_p, _c_p, _i64, _i32, _f32, _i32, _c_p, _c_p],
    "rvb_vat_finalize_stats": [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _i64, _i32, _f32, _f32, _f32, _i32, _c_p, _c_p, _c_p,
                               _i64, _c_p],
    "rvb_bce_mean": [_c_p, _c_p, _i64, _c_p, _c_p, _c_p],
    "rvb_bn_reduce": [_c_p, _c_p, _c_p, _i32, _i32, _i64, _i32, _c_p, _c_p],
    "rvb_bn_forward": [_c_p, _i32, _i32, _i64, _i32, _c_p, _c_p, _c_p, _f32, _f32, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p],
    "rvb_bn_train_forward": [_c_p, _i32, _i32, _i64, _c_p, _c_p, _f32, _f32, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p],
    "rvb_bn_train_backward": [_c_p, _c_p, _i32, _i32, _i64, _c_p, _c_p, _c_p, _i32, _c_p, _c_p, _c_p, _c_p, _c_p],
    "rvb_bn_train_forward_nhwc": [_c_p, _i64, _i32, _c_p, _c_p, _f32, _f32, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p],
    "rvb_bn_apply_nhwc": [_c_p, _i64, _i32, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p],
    "rvb_bn_train_backward_nhwc": [_c_p, _c_p, _i64, _i32, _c_p, _c_p, _c_p, _i32, _c_p, _c_p, _c_p, _c_p, _c_p],
    "rvb_bn_apply": [_c_p, _i32, _i32, _i64, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p],
    "rvb_bn_backward": [_c_p, _c_p, _i32, _i32, _i64, _i32, _c_p, _c_p, _c_p, _c_p, _i32, _c_p, _c_p, _c_p, _c_p],
}
ABI_VERSION = 1
BCE_WORKSPACE_FLOATS = 1032

DIV_BCE, DIV_BKL, DIV_MSE = 0, 1, 2
PAD_REFLECT, PAD_CONSTANT, PAD_NONE = 0, 1, 2
EPI_POWER, EPI_MAGNITUDE, EPI_COMPLEX, EPI_PHASE, EPI_POWER_P = 0, 1, 2, 3, 4
EPI_TIME_MAJOR = 0x10
LAYOUT_BINS_MAJOR, LAYOUT_TIME_MAJOR = 0, 1

_lib = None


class RvbError(RuntimeError):
    pass


def load():
    """dlopen librvb.so (works without a GPU: the CUDA driver is only touched by the first launch)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "reconvat_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C reconvat_b200/csrc`. There is no CPU / PyTorch fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.rvb_abi_version.restype = ctypes.c_int
    lib.rvb_last_error.restype = ctypes.c_char_p
    lib.rvb_launch_count.restype = ctypes.c_int64
    lib.rvb_vat_stats_workspace_bytes.restype = ctypes.c_int64
    lib.rvb_vat_stats_workspace_bytes.argtypes = [_i64]
    lib.rvb_parity_plane_len.restype = ctypes.c_int64
    lib.rvb_parity_plane_len.argtypes = [_i32, _i32, _i32, _i32, _i32, _i32]
    lib.rvb_bn_splits.restype = ctypes.c_int
    lib.rvb_bn_nhwc_workspace_bytes.restype = ctypes.c_int64
    lib.rvb_bn_nhwc_workspace_bytes.argtypes = [_i32]
    lib.rvb_bn_splits.argtypes = [_i32, _i32, _i64]
    if lib.rvb_abi_version() != ABI_VERSION:
        raise ImportError("reconvat_b200: librvb.so has ABI %d, the Python side expects %d -- rebuild"
                          % (lib.rvb_abi_version(), ABI_VERSION))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    _lib = lib
    return lib


def launch_count():
    """Kernels launched by librvb.so in this process so far (bench.py reports the delta)."""
    return int(load().rvb_launch_count())


def parity_plane_len(n_samples, pad, pad_mode, n_fft, hop, n_frames):
    return int(load().rvb_parity_plane_len(int(n_samples), int(pad), int(pad_mode), int(n_fft), int(hop), int(n_frames)))


def bn_nhwc_workspace_bytes(c):
    return int(load().rvb_bn_nhwc_workspace_bytes(int(c)))


def bn_splits(n, c, hw):
    return int(load().rvb_bn_splits(int(n), int(c), int(hw)))


def vat_stats_workspace_bytes(n_rows):
    return int(load().rvb_vat_stats_workspace_bytes(int(n_rows)))


def _stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t, dtype=torch.float32):
    """Device pointer of a CUDA tensor; refuses anything the kernels cannot read."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RvbError("reconvat_b200 kernels need CUDA tensors (got %s); there is no CPU path"
                       % (t.device if isinstance(t, torch.Tensor) else type(t)))
    if t.dtype != dtype:
        raise RvbError("expected %s, got %s" % (dtype, t.dtype))
    return t.data_ptr()


# Optional per-entry-point CUDA-event timing (bench.py's roofline numbers): {name: [(start, end), ...]}
_event_log = None


def record_events(names):
    """Start recording (start, end) CUDA events around every call of the given entry points on the
    launching stream; returns the log dict.  ``record_events(None)`` stops recording."""
    global _event_log
    _event_log = None if names is None else {n: [] for n in names}
    return _event_log


# Optional call recording (bench.py re-launches single entry points back to back): [(name, args), ...]
_call_log = None


def record_calls(log):
    """Append (name, args) of every entry-point call to ``log`` (a list) until ``record_calls(None)``.  The arguments
    are raw pointers: whoever replays them with :func:`raw_call` keeps the memory they point to alive."""
    global _call_log
    _call_log = log


def raw_call(name, args):
    """Re-issue a recorded call on the current stream."""
    rc = getattr(load(), name)(*args, _stream())
    if rc != 0:
        raise RvbError("%s failed (%d): %s" % (name, rc, load().rvb_last_error().decode("utf-8", "replace")))


def call_on(name, stream, *args):
    """``call`` for callers that already hold the stream handle (saves one stream query per call)."""
    lib = _lib if _lib is not None else load()
    if _call_log is not None or _event_log is not None:
        return call(name, *args)
    rc = getattr(lib, name)(*args, stream)
    if rc != 0:
        raise RvbError("%s failed (%d): %s" % (name, rc, lib.rvb_last_error().decode("utf-8", "replace")))


def call(name, *args):
    """Invoke an entry point on the current PyTorch stream; raise RvbError on a non-zero status."""
    lib = load()
    if _call_log is not None:
        _call_log.append((name, args))
    log = _event_log.get(name) if _event_log is not None else None
    if log is not None:
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
    rc = getattr(lib, name)(*args, _stream())
    if log is not None:
        end.record()
        log.append((start, end))
    if rc != 0:
        raise RvbError("%s failed (%d): %s" % (name, rc, lib.rvb_last_error().decode("utf-8", "replace")))
