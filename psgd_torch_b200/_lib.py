"""ctypes binding of libpsgd_b200.so (include/psgd_b200.h).

The product path has exactly one implementation: the CUDA library.  If the shared object is missing or the
device is not a B200-class GPU (compute capability 10.x) every entry point raises -- there is no CPU or
PyTorch fallback (and nothing under oracle/ is ever imported from here).
"""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpsgd_b200.so")

PSGD_BF16, PSGD_F32 = 0, 1
PSGD_DIAG, PSGD_DENSE = 0, 1
MAX_BATCH = 16   # units per batched call (KB_MAX in csrc/kron_kernels.cuh)
# psgd_dq_t / stage bits of psgd_kron_update
DQ_CODES = {"Q0.5EQ1.5": 0, "Q0p5EQ1p5": 0, "EQ": 1, "QEP": 2, "QEQ": 3, "QUAD": 4, "QUAD4P": 5, "PRO4P": 6}
STAGE_PREPARE, STAGE_FACTOR_L, STAGE_FACTOR_R, STAGE_BALANCE = 1, 2, 4, 8

_DTYPES = {torch.bfloat16: PSGD_BF16, torch.float32: PSGD_F32}


class EngineError(RuntimeError):
    pass


class KronT(C.Structure):
    _fields_ = [("m", C.c_int32), ("n", C.c_int32), ("kind_l", C.c_int32), ("kind_r", C.c_int32), ("dtype", C.c_int32),
                ("has_r", C.c_int32), ("QL", C.c_void_p), ("QR", C.c_void_p), ("LL", C.c_void_p), ("LR", C.c_void_p)]


class KronNoiseT(C.Structure):
    _fields_ = [("N", C.c_void_p), ("V0_spd_l", C.c_void_p), ("V0_skh_l", C.c_void_p), ("V0_spd_r", C.c_void_p),
                ("V0_skh_r", C.c_void_p), ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64)]


class LraT(C.Structure):
    _fields_ = [("n", C.c_int64), ("r", C.c_int32), ("dtype", C.c_int32), ("U", C.c_void_p), ("V", C.c_void_p),
                ("d", C.c_void_p), ("Lu", C.c_void_p), ("Lv", C.c_void_p), ("Ld", C.c_void_p)]


# every symbol include/psgd_b200.h declares: name -> (restype, argtypes)
_vp, _i, _f, _sz, _i64 = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_int64
SYMBOLS = {
    "psgd_abi_version": (_i, []),
    "psgd_status_string": (C.c_char_p, [_i]),
    "psgd_last_error": (C.c_char_p, [_vp]),
    "psgd_create": (_i, [C.POINTER(_vp), _i]),
    "psgd_destroy": (None, [_vp]),
    "psgd_set_gemm_path": (_i, [_vp, _i]),
    "psgd_launch_count": (_i64, [_vp]),
    "psgd_set_sm_limit": (_i, [_vp, _i]),
    "psgd_peer_enable": (_i, [_i]),
    "psgd_peer_copy_async": (_i, [_vp, _vp, _sz, _vp]),
    "psgd_set_fp32_tensor_cores": (_i, [_vp, _i]),
    "psgd_kron_workspace_bytes": (_sz, [_vp, C.POINTER(KronT)]),
    "psgd_kron_whiten_q0p5eq1p5_update": (_i, [_vp, C.POINTER(KronT), _vp, _f, _f, _f, C.POINTER(KronNoiseT), _i, _vp, _sz, _vp]),
    "psgd_kron_precond_grad": (_i, [_vp, C.POINTER(KronT), _vp, _vp, _vp, _vp, _sz, _vp]),
    "psgd_kron_batch_workspace_bytes": (_sz, [_vp, C.POINTER(KronT), _i]),
    "psgd_kron_whiten_q0p5eq1p5_update_batched": (_i, [_vp, C.POINTER(KronT), _i, C.POINTER(_vp), _f, _f, _f, C.POINTER(KronNoiseT),
                                                  C.POINTER(_i), _vp, _sz, _vp]),
    "psgd_kron_precond_grad_batched": (_i, [_vp, C.POINTER(KronT), _i, C.POINTER(_vp), C.POINTER(_vp), _vp, _vp, _sz, _vp]),
    "psgd_kron_balance": (_i, [_vp, C.POINTER(KronT), _vp, _sz, _vp]),
    "psgd_kron_update_workspace_bytes": (_sz, [_vp, C.POINTER(KronT), _i]),
    "psgd_kron_update": (_i, [_vp, C.POINTER(KronT), _i, _vp, _vp, _f, _f, _f, C.POINTER(KronNoiseT), _i, _vp, _sz, _vp]),
    "psgd_kron_apply_factors": (_i, [_vp, C.POINTER(KronT), _vp, _vp, _vp, _vp, _sz, _vp]),
    "psgd_kron_solve_factors": (_i, [_vp, C.POINTER(KronT), _vp, _vp, _vp, _sz, _vp]),
    "psgd_kron_factor_step": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "psgd_procrustes_step3": (_i, [_vp, _i, _vp, _i, _vp, _f, _vp, _sz, _vp]),
    "psgd_symmetry_gap": (_i, [_vp, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "psgd_helper_workspace_bytes": (_sz, [_vp, _i, _i]),
    "psgd_norm_lower_bound_spd": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "psgd_norm_lower_bound_skh": (_i, [_vp, _i, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "psgd_procrustes_step2": (_i, [_vp, _i, _vp, _i, _vp, _f, _vp, _sz, _vp]),
    "psgd_kron_factor_update": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _f, _f, _f, _vp, _vp, _vp, _sz, _vp]),
    "psgd_kwns4_head": (_i, [_vp, _i64, _vp, _i, _vp, _i, _f, _f, _i, _vp, _vp, _i, _f, _vp]),
    "psgd_kwns4_tail": (_i, [_vp, _i64, _i64, _vp, _i, _vp, _i, _vp, _f, _f, _f, _vp]),
    "psgd_lra_workspace_bytes": (_sz, [_vp, C.POINTER(LraT)]),
    "psgd_lra_update": (_i, [_vp, C.POINTER(LraT), _vp, _vp, _f, _f, _i, _vp, _sz, _vp]),
    "psgd_lra_whiten_update": (_i, [_vp, C.POINTER(LraT), _vp, _vp, _f, _f, _f, _i, _vp, _sz, _vp]),
    "psgd_lra_newton_update": (_i, [_vp, C.POINTER(LraT), _vp, _vp, _vp, _f, _f, _f, _i, _vp, _sz, _vp]),
    "psgd_lra_precond_grad": (_i, [_vp, C.POINTER(LraT), _vp, _vp, _vp, _vp, _sz, _vp]),
    "psgd_lra_workspace_offsets": (_i, [_vp, C.POINTER(LraT), C.POINTER(_sz), C.POINTER(_sz)]),
    "psgd_lra_update_staged": (_i, [_vp, C.POINTER(LraT), _vp, _vp, _f, _f, _f, _i, _i, _i, _vp, _sz, _vp]),
    "psgd_lra_precond_grad_staged": (_i, [_vp, C.POINTER(LraT), _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "psgd_gemm": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _f, _vp, _i, _f, _vp]),
    "psgd_timing_enable": (_i, [_vp, _i]),
    "psgd_timing_read": (_i, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "psgd_timing_gemm_launches": (_i64, [_vp]),
    "psgd_timing_executed_flops": (C.c_double, [_vp]),
    "psgd_debug_set_flags": (_i, [_vp, _i]),
    "psgd_debug_set_tile_n": (_i, [_vp, _i]),
    "psgd_debug_set_mn_desc": (_i, [_vp, _i, _i]),
    "psgd_debug_read_probe": (_i, [_vp, _vp, _sz, _i, _i, _i, _i, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()
_handles = {}
_workspaces = {}


def load_library():
    """dlopen the in-tree shared object (works without a GPU: the CUDA runtime is linked statically and the driver
    entry points are resolved lazily in psgd_create)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise EngineError(f"{LIB_PATH} not found: build it with `python -m psgd_torch_b200.build` "
                                  "(there is no fallback implementation)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
                fn.restype = res
                fn.argtypes = args
            if lib.psgd_abi_version() != 2:
                raise EngineError("libpsgd_b200.so ABI version mismatch")
            _lib = lib
    return _lib


def check(handle, rc, what):
    if rc != 0:
        lib = load_library()
        msg = lib.psgd_status_string(rc).decode()
        extra = lib.psgd_last_error(handle).decode() if handle else ""
        raise EngineError(f"{what} failed: {msg} ({rc}) {extra}")


def _dev_index(device):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise EngineError(f"psgd_torch_b200 runs on CUDA (sm_100a) tensors only, got a tensor on '{dev}'")
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        # kernels launch on the process's current device: a tensor of another device would be handed to the wrong GPU
        raise EngineError(f"tensor lives on cuda:{idx} but the current device is cuda:{torch.cuda.current_device()}: wrap the call in "
                          f"`with torch.cuda.device({idx}):` (one process per GPU is the supported layout)")
    return idx


def handle_for(device):
    """One engine context per (CUDA device, current stream).  A context owns the split-K partial-tile buffer and arrival counters of the
    tcgen05 GEMM, which assume stream order between the launches that share them, so calls issued on different streams get different
    contexts (42 MB each) instead of racing on one."""
    lib = load_library()
    idx = _dev_index(device)
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    h = _handles.get(key)
    if h is None:
        with _lock:
            h = _handles.get(key)
            if h is None:
                out = C.c_void_p()
                rc = lib.psgd_create(C.byref(out), idx)
                check(None, rc, f"psgd_create(device={idx})")
                h = out
                if idx in _sm_limit:
                    check(h, lib.psgd_set_sm_limit(h, _sm_limit[idx]), "psgd_set_sm_limit")
                _handles[key] = h
    return h


def workspace(device, nbytes):
    """A scratch tensor per (device, current stream), grown geometrically: engine calls on one stream are stream-ordered, so they can
    share one buffer; calls on different streams never do (same keying as handle_for)."""
    idx = _dev_index(device)
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = None
        _workspaces[key] = None
        ws = torch.empty(max(int(nbytes * 1.25), 1 << 20), dtype=torch.uint8, device=torch.device("cuda", idx))
        _workspaces[key] = ws
    return ws


def free_workspaces():
    _workspaces.clear()


def dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise EngineError(f"unsupported preconditioner dtype {t.dtype}; the engine computes in bfloat16 or float32")


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def stream_ptr(device):
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    return C.c_void_p(torch.cuda.current_stream(idx).cuda_stream)


def set_fp32_tensor_cores(on, device=None):
    """Opt-in: big fp32 products on the tensor cores as bf16 triples (include/psgd_b200.h: psgd_set_fp32_tensor_cores)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    h = handle_for(dev)
    check(h, load_library().psgd_set_fp32_tensor_cores(h, int(bool(on))), "psgd_set_fp32_tensor_cores")


_sm_limit = {}


def set_sm_limit(sms, device=None):
    """Leave SMs free for concurrent kernels of other streams (NCCL): the engine's persistent kernels size their grids for `sms` SMs
    (<= 0: all).  Applies to every context of the device, present and future (include/psgd_b200.h: psgd_set_sm_limit)."""
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = _dev_index(dev)
    _sm_limit[idx] = int(sms)
    lib = load_library()
    for (i, _), h in _handles.items():
        if i == idx:
            check(h, lib.psgd_set_sm_limit(h, int(sms)), "psgd_set_sm_limit")


def launch_count(device=None):
    """Kernels launched so far by every context of `device` (all devices if None)."""
    lib = load_library()
    if device is None:
        return sum(lib.psgd_launch_count(h) for h in _handles.values())
    idx = _dev_index(device)
    handle_for(device)
    return sum(lib.psgd_launch_count(h) for (i, _), h in _handles.items() if i == idx)
