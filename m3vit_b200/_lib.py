"""ctypes binding of the C-ABI library (include/m3vit_moe.h).

The library is the product: there is NO fallback.  If the shared object is
missing or the device is not a B200 the import / first call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# M3_LIB_PATH: a diagnosis build of the SAME library (e.g. lib/libm3vit_moe_trace.so, M3_GEMM_TRACE=1 python -m m3vit_b200.build)
LIB_PATH = os.environ.get("M3_LIB_PATH") or os.path.join(HERE, "lib", "libm3vit_moe.so")

M3_F32, M3_BF16 = 0, 1
PAD_ROWS = 256     # expert queues are padded to the 256-row CTA-pair tile of the tensor-core GEMM

_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/m3vit_moe.h one to one
SIGNATURES = {
    "m3_abi_version": (_i, []),
    "m3_status_string": (C.c_char_p, [_i]),
    "m3_check_device": (_i, []),
    "m3_gate_num_partials": (_i, [_i, _i]),
    "m3_gate_fwd": (_i, [_p, _i, _i64, _p, _p, _p, _f, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "m3_gate_fwd_rng": (_i, [_p, _i, _i64, _p, _p, _p, _f, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "m3_gate_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "m3_gate_bwd": (_i, [_p, _i, _i64, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                         _p, _p, _p, _sz, _p]),
    "m3_route_plan_workspace_bytes": (_sz, [_i, _i, _i]),
    "m3_route_max_rows": (_i, [_i, _i, _i, _i]),
    "m3_route_max_tiles": (_i, [_i, _i, _i, _i]),
    "m3_route_plan": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "m3_dispatch_fwd": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _p, _i, _p]),
    "m3_dispatch_bwd": (_i, [_p, _i, _p, _i, _i, _i, _p, _p, _i, _p, _i, _p]),
    "m3_combine_fwd": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _i, _p]),
    "m3_combine_bwd": (_i, [_p, _i, _p, _i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _p, _p]),
    "m3_ffn_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "m3_ffn_saved_bytes": (_sz, [_i, _i, _i, _i]),
    "m3_ffn_uses_chain": (_i, [_i, _i, _i]),
    "m3_set_gemm_sm_limit": (_i, [_i]),
    "m3_set_knob": (_i, [_i, _i]),
    "m3_ffn_fwd": (_i, [_i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "m3_ffn_fwd_dropout": (_i, [_i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _sz, _f, _p, _p]),
    "m3_ffn_bwd": (_i, [_i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "m3_ffn_bwd_parts": (_i, [_i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _f,
                              _p, _i, _p]),
    "m3_ffn_bwd_dropout": (_i, [_i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _f,
                                _p, _p]),
    "m3_cast_weights_bf16": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "m3_ep_plan": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "m3_ep_dispatch_fwd": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _i, _p]),
    "m3_ep_ffn_fwd": (_i, [_i, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _f, _p, _p]),
    "m3_ep_ffn_bwd": (_i, [_i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _f,
                           _p, _i, _p]),
    "m3_ep_combine_fwd": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _p, _i, _p, _p]),
    "m3_ep_combine_bwd": (_i, [_p, _i, _p, _p, _i, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "m3_ep_dispatch_bwd": (_i, [_p, _i, _p, _p, _i, _i, _i, _p, _p, _i, _p, _i, _p]),
    "m3_zero_pad_rows": (_i, [_p, _i, _p, _p, _i, _i, _p, _p]),
    "m3_ep_barrier": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "m3_ln_stats": (_i, [_p, _i, _i, _f, _p, _p, _p]),
    "m3_ln_fold_gate": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "m3_gate_fwd_ln": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p, _f, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p,
                            _p, _p]),
    "m3_dispatch_fwd_ln": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _i, _p]),
    "m3_combine_fwd_res": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _p, _p]),
    "m3_gate_bwd_ln": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p,
                            _p, _p, _p, _p, _p, _sz, _p]),
    "m3_ln_bwd_workspace_bytes": (_sz, [_i, _i]),
    "m3_ln_bwd_res": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "m3_debug_trace_buffer": (_i, [_p, _i]),
    "m3_ipc_alloc": (_i, [_sz, C.POINTER(_p), _p]),
    "m3_ipc_open": (_i, [_p, C.POINTER(_p)]),
    "m3_ipc_close": (_i, [_p]),
    "m3_ipc_free": (_i, [_p]),
}

_lib = None


class M3Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """dlopen the C-ABI library and bind every declared symbol.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise M3Error(
            f"{LIB_PATH} not found: the CUDA extension is the product and there is no fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError => header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.m3_abi_version() != 2:
        raise M3Error("libm3vit_moe.so ABI version mismatch")
    # tuning knobs from the environment, e.g. M3_KNOBS="0=1,3=2" (m3_set_knob(knob, value); include/m3vit_moe.h)
    for kv in filter(None, os.environ.get("M3_KNOBS", "").split(",")):
        k, v = kv.split("=")
        if lib.m3_set_knob(int(k), int(v, 0)) < 0:
            raise M3Error(f"M3_KNOBS: unknown knob {k}")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status == 0:
        return
    msg = load().m3_status_string(status).decode()
    if status < 0:
        raise ValueError(f"{what}: {msg} (m3 status {status})")
    raise M3Error(f"{what}: CUDA error {status}: {msg}")


_device_ok = {}


def require_device(t: torch.Tensor) -> None:
    """Every product entry point starts here: CUDA tensor on an sm_100 device."""
    if not t.is_cuda:
        raise M3Error("m3vit_b200 runs on B200 GPUs only: got a CPU tensor (there is no CPU fallback)")
    dev = t.device.index
    if dev != torch.cuda.current_device():
        raise M3Error(f"tensor on cuda:{dev} but the current device is cuda:{torch.cuda.current_device()}: the kernels launch "
                      "on the current device's stream (use torch.cuda.set_device / torch.cuda.device(...))")
    if dev not in _device_ok:
        with torch.cuda.device(dev):
            check(load().m3_check_device(), "m3_check_device")
        _device_ok[dev] = True


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return M3_F32
    if t.dtype == torch.bfloat16:
        return M3_BF16
    raise ValueError(f"unsupported dtype {t.dtype} (fp32 or bf16)")


def ptr(t):
    return None if t is None else t.data_ptr()


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device.  The raw C accessor is ~50x cheaper than
    torch.cuda.current_stream() (16 us of Python per call: it showed up as a quarter of the host time per layer call).
    Kernels are launched on the CURRENT device: require_device() refuses tensors that live on another one."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream
