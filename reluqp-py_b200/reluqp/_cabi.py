"""ctypes binding of librqp.so (include/rqp.h).  PyTorch tensors are the only glue: every
pointer handed to the library is ``tensor.data_ptr()`` of a tensor the caller keeps alive.

There is NO fallback: if the library is missing or the device is not a CUDA device the
solve path raises.  Build the library with ``make -C reluqp-py_b200`` (or
``python __graft_entry__.py``)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "lib", "librqp.so")

RQP_F32, RQP_F64 = 0, 1
RQP_OK = 0
RQP_STATUS_SOLVED, RQP_STATUS_MAX_ITER, RQP_STATUS_RUNNING = 0, 1, 2
RQP_ERR_WATCHDOG = -6
RQP_TRACE_STRIDE = 5
EPOCH_LIMIT = 0x70000000

EXPORTS = ("rqp_query", "rqp_size_limit", "rqp_workspace_size", "rqp_solve", "rqp_structured_workspace_size",
           "rqp_solve_structured", "rqp_update_bias", "rqp_resolve",
           "rqp_batch_workspace_size", "rqp_solve_batched", "rqp_copy_h2d", "rqp_stream_sync",
           "rqp_probe_bandwidth", "rqp_kernel_launches",
           "rqp_strerror", "rqp_last_cuda_error")


class rqp_caps(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("sm_count", C.c_int32), ("max_smem_per_block", C.c_int32),
                ("cooperative_launch", C.c_int32), ("l2_bytes", C.c_int64),
                ("global_mem_bytes", C.c_int64), ("max_clusters8", C.c_int32), ("cl_min_cell_bytes", C.c_int32)]


class rqp_problem(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("nx", C.c_int32), ("nc", C.c_int32), ("n_rho", C.c_int32),
                ("ldw", C.c_int64),
                ("W", C.c_void_p), ("b", C.c_void_p), ("H", C.c_void_p), ("A", C.c_void_p),
                ("AT", C.c_void_p), ("g", C.c_void_p), ("l", C.c_void_p), ("u", C.c_void_p),
                ("rhos", C.c_void_p)]


class rqp_structured(C.Structure):
    _fields_ = [("M", C.c_void_p), ("Rv", C.c_void_p), ("Rinv", C.c_void_p), ("Apad", C.c_void_p),
                ("ldm", C.c_int64), ("lda", C.c_int64)]


class rqp_settings(C.Structure):
    _fields_ = [("max_iter", C.c_int32), ("check_interval", C.c_int32), ("adaptive_rho", C.c_int32),
                ("poll_backoff_ns", C.c_int32),
                ("eps_abs", C.c_double), ("eps_rel", C.c_double), ("rho_min", C.c_double),
                ("rho_max", C.c_double), ("adaptive_rho_tolerance", C.c_double),
                ("grid", C.c_int32), ("block", C.c_int32), ("w_residency", C.c_int32),
                ("watchdog_ms", C.c_int32), ("prepoll_cycles", C.c_int32), ("exchange_flags", C.c_int32)]


class rqp_state(C.Structure):
    _fields_ = [("v", C.c_void_p), ("rho_ind", C.c_int32), ("epoch", C.c_uint32), ("x_host", C.c_void_p),
                ("post_seq", C.c_uint64)]


class rqp_result(C.Structure):
    _fields_ = [("iter", C.c_int32), ("status", C.c_int32), ("rho_ind", C.c_int32), ("error", C.c_int32),
                ("pri_res", C.c_double), ("dua_res", C.c_double), ("rho_estimate", C.c_double),
                ("obj_val", C.c_double),
                ("n_checks", C.c_int32), ("n_rho_switches", C.c_int32),
                ("t_begin_ns", C.c_uint64), ("t_end_ns", C.c_uint64),
                ("grid", C.c_int32), ("block", C.c_int32), ("rows_per_cta", C.c_int32),
                ("rows_in_smem", C.c_int32), ("phase_cycles", C.c_uint64 * 8), ("seq", C.c_uint64)]


class rqp_batch(C.Structure):
    _fields_ = [("B", C.c_int32), ("ldv", C.c_int32),
                ("V", C.c_void_p), ("L", C.c_void_p), ("U", C.c_void_p), ("G", C.c_void_p),
                ("Bmat", C.c_void_p),
                ("rho_ind", C.c_void_p), ("iter", C.c_void_p), ("status", C.c_void_p),
                ("pri_res", C.c_void_p), ("dua_res", C.c_void_p), ("rho_estimate", C.c_void_p),
                ("engine", C.c_int32), ("res_planes", C.c_int32), ("W_hi", C.c_void_p), ("W_lo", C.c_void_p),
                ("reserved_dbg", C.c_void_p), ("kmask", C.c_void_p), ("kmask_min_blocks", C.c_int32),
                ("first_window_ms", C.POINTER(C.c_float)),
                ("reduced", C.c_int32), ("Wr", C.c_void_p), ("br", C.c_void_p), ("Bred", C.c_void_p),
                ("Rv", C.c_void_p), ("Rinv", C.c_void_p)]


_lib = None


def load():
    """Load librqp.so once; raise loudly if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "librqp.so not found at {}: the ReLU-QP solve path has no CPU or PyTorch fallback. "
            "Build it with `make -C reluqp-py_b200` (needs nvcc, sm_100a).".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    vp, i32, sz = C.c_void_p, C.c_int32, C.c_size_t
    lib.rqp_query.argtypes = [C.c_int, C.POINTER(rqp_caps)]
    lib.rqp_size_limit.argtypes = [i32, C.POINTER(i32)]
    lib.rqp_workspace_size.argtypes = [C.POINTER(rqp_problem), C.POINTER(rqp_settings), C.POINTER(sz)]
    lib.rqp_solve.argtypes = [C.POINTER(rqp_problem), C.POINTER(rqp_settings), C.POINTER(rqp_state),
                              vp, vp, i32, vp, sz, vp]
    lib.rqp_structured_workspace_size.argtypes = [C.POINTER(rqp_problem), C.POINTER(rqp_structured),
                                                  C.POINTER(rqp_settings), C.POINTER(sz)]
    lib.rqp_solve_structured.argtypes = [C.POINTER(rqp_problem), C.POINTER(rqp_structured), C.POINTER(rqp_settings),
                                         C.POINTER(rqp_state), vp, vp, i32, vp, sz, vp]
    lib.rqp_update_bias.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp]
    lib.rqp_resolve.argtypes = [C.POINTER(rqp_problem), C.POINTER(rqp_settings), C.POINTER(rqp_state),
                                vp, vp, i32, vp, sz, vp, vp, sz, i32, vp, vp, sz, vp]
    lib.rqp_batch_workspace_size.argtypes = [C.POINTER(rqp_problem), C.POINTER(rqp_settings), i32, C.POINTER(sz)]
    lib.rqp_solve_batched.argtypes = [C.POINTER(rqp_problem), C.POINTER(rqp_settings), C.POINTER(rqp_batch),
                                      vp, sz, C.POINTER(i32), vp]
    lib.rqp_probe_bandwidth.argtypes = [vp, sz, i32, C.POINTER(C.c_float), vp]
    lib.rqp_copy_h2d.argtypes = [vp, vp, sz, vp]
    lib.rqp_stream_sync.argtypes = [vp]
    for name in EXPORTS:
        getattr(lib, name).restype = C.c_int
    lib.rqp_kernel_launches.argtypes = []
    lib.rqp_kernel_launches.restype = C.c_ulonglong
    lib.rqp_strerror.argtypes = [C.c_int]
    lib.rqp_strerror.restype = C.c_char_p
    lib.rqp_last_cuda_error.argtypes = []
    lib.rqp_last_cuda_error.restype = C.c_char_p
    _lib = lib
    return lib


def raw_stream(device_index):
    """cudaStream_t of torch's current stream on that device, as an int (the cheap way when torch has it)."""
    import torch
    try:
        return torch._C._cuda_getCurrentRawStream(device_index)
    except AttributeError:                       # pragma: no cover - older / newer torch layouts
        return torch.cuda.current_stream(device_index).cuda_stream


def check(rc, what):
    if rc != RQP_OK:
        lib = load()
        msg = lib.rqp_strerror(rc).decode()
        if rc == -3:
            msg += " ({})".format(lib.rqp_last_cuda_error().decode())
        raise RuntimeError("{} failed: {} [{}]".format(what, msg, rc))


def query(device_index=0):
    caps = rqp_caps()
    check(load().rqp_query(int(device_index), C.byref(caps)), "rqp_query")
    return caps


def dtype_code(torch_dtype):
    import torch
    if torch_dtype == torch.float64:
        return RQP_F64
    if torch_dtype == torch.float32:
        return RQP_F32
    raise ValueError("ReLU-QP supports float32 and float64, got {}".format(torch_dtype))
