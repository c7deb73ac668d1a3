"""ctypes binding of ``libszn.so`` (C ABI declared in ``include/szn.h``).

There is no CPU fallback: if the shared library is missing or a call fails, a ``RuntimeError`` is
raised.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C
zeroshotsemanticsegmentation_b200/csrc``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SZN_LIB") or os.path.join(_HERE, "libszn.so")  # SZN_LIB: A/B-test another build

I, LL, P, ULL = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_ulonglong

# name -> argtypes (must mirror include/szn.h; tests/test_abi.py checks every symbol is exported)
SIGNATURES = {
    "szn_conv_fwd": [I, P, P, P, P, I, I, I, I, I, I, I, I, I, P, I, I, LL, P],
    "szn_conv_dgrad": [I, P, P, P, I, I, I, I, I, I, I, I, P, P, I, LL, P, P],
    "szn_conv_wgrad": [I, P, P, P, I, I, I, I, I, I, I, I, LL, P],
    "szn_set_wgrad_waves": [I],
    "szn_comm_unique_id": [P],
    "szn_comm_init": [P, I, I, P],
    "szn_comm_destroy": [P],
    "szn_allreduce_bucket": [P, P, P, I, I, P],
    "szn_conv1_1_fwd": [I, P, P, P, P, I, I, I, I, P],
    "szn_conv1_1_wgrad": [I, P, P, P, I, I, I, I, P],
    "szn_pool_fwd": [I, P, P, I, I, I, I, P],
    "szn_pool_bwd": [I, P, P, P, I, I, I, I, I, P, P],
    "szn_pool_fwd_code": [I, P, P, P, I, I, I, I, P],
    "szn_pool_bwd_code": [I, P, P, P, I, I, I, I, I, P, P],
    "szn_bias_grad": [I, P, P, LL, I, LL, P],
    "szn_pack_weight": [I, P, P, I, I, I, I, I, P],
    "szn_pack_weight_dgrad": [I, P, P, I, I, I, I, I, I, P],
    "szn_col2im": [I, P, P, I, I, I, I, I, I, P],
    "szn_unpack_wgrad": [P, P, I, I, I, I, P],
    "szn_cast": [I, P, P, LL, I, P],
    "szn_dropout_scale": [P, I, ULL, P],
    "szn_upsample32_crop_fwd": [P, P, I, I, I, I, I, I, I, I, P],
    "szn_upsample32_crop_bwd": [I, P, P, I, I, I, I, I, I, I, I, P],
    "szn_deconv_small_fwd": [P, P, P, I, I, I, I, I, I, I, I, I, P],
    "szn_deconv_small_dgrad": [I, P, P, P, I, I, I, I, I, I, I, I, I, P],
    "szn_deconv_small_wgrad": [P, P, P, I, I, I, I, I, I, I, I, I, P],
    "szn_embed_loss_fwd": [I, P, P, P, P, I, I, I, I, I, P, P, P, P],
    "szn_embed_loss_bwd": [I, P, P, P, P, I, I, I, I, I, P, P, P, P, P],
    "szn_ce2d_fwd": [P, P, I, I, I, I, I, P, P, P, P],
    "szn_ce2d_bwd": [P, P, I, I, I, I, I, P, P, P, P, P],
    "szn_loss_finalize": [I, P, P, P],
    "szn_embed_argmax": [P, P, I, I, I, I, I, P, P, P],
    "szn_stitch_labels": [P, P, P, P, P, I, I, I, I, P, P],
    "szn_confusion_hist": [P, P, LL, I, P, P, P],
    "szn_head_fused_fwd": [I, P, I, I, P, P, I, I, I, I, I, I, I, P, P, P, P, P],
    "szn_head_fused_bwd": [I, P, I, I, P, I, I, I, I, I, P, P, P, P, P],
    "szn_sgd_step": [P, P, P, LL, ctypes.c_float, ctypes.c_float, ctypes.c_float, I, P],
    "szn_adam_step": [P, P, P, P, LL] + [ctypes.c_float] * 6 + [P],
}

F32, BF16, F32X3 = 0, 1, 2

_lib = None


def load():
    """Load libszn.so once; raises RuntimeError (never falls back) when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "libszn.so is not built (%s): run `make -C zeroshotsemanticsegmentation_b200/csrc` "
            "or __graft_entry__.build(); there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = I
    lib.szn_last_error.restype = ctypes.c_char_p
    lib.szn_launch_count.restype = LL
    lib.szn_embed_argmax_scratch_floats.argtypes = [I, I]
    lib.szn_embed_argmax_scratch_floats.restype = LL
    lib.szn_abi_version.restype = I
    lib.szn_comm_available.restype = I
    lib.szn_head_fused_workspace_floats.argtypes = [I, I, I, I]
    lib.szn_head_fused_workspace_floats.restype = LL
    lib.szn_build_id.argtypes = []
    lib.szn_build_id.restype = ctypes.c_char_p
    _lib = lib
    return lib


def build_id():
    """Hash of the kernel sources and compiler flags this libszn.so was built from (include/szn_build.h)."""
    return load().szn_build_id().decode()


_profiler = None


def set_profiler(fn):
    """``fn(name, args) -> done()`` brackets every C-ABI call (bench.py's per-kernel CUDA-event timing); None disables."""
    global _profiler
    _profiler = fn


NVTX = os.environ.get("SZN_NVTX") == "1"  # SZN_NVTX=1: an NVTX range per layer (engine.py) and per C-ABI call


class nvtx_range:
    """``with nvtx_range("conv3_2 fwd"):`` -- a no-op unless SZN_NVTX=1 (nsys / ncu --nvtx show the layer structure)."""

    def __init__(self, label):
        self.label = label

    def __enter__(self):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_push(self.label)

    def __exit__(self, *exc):
        if NVTX:
            import torch
            torch.cuda.nvtx.range_pop()


def call(name, *args):
    lib = load()
    done = _profiler(name, args) if _profiler is not None else None
    if NVTX:
        with nvtx_range(name):
            rc = getattr(lib, name)(*args)
    else:
        rc = getattr(lib, name)(*args)
    if done is not None:
        done()
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, lib.szn_last_error().decode()))


def launch_count():
    return int(load().szn_launch_count())


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream
