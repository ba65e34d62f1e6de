"""ctypes binding of navierstokes_b200/lib/libnsk.so -- the C ABI declared in include/nsk.h.

There is no fallback: if the library is missing, or no sm_100 GPU is present when a context is
created, the call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libnsk.so"

c_void_pp = C.POINTER(C.c_void_p)
c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_int64_p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); must list every symbol include/nsk.h declares (tests check this)
SIGNATURES = {
    "nsk_version": (C.c_int, []),
    "nsk_strerror": (C.c_char_p, [C.c_int]),
    "nsk_last_error": (C.c_char_p, [C.c_void_p]),
    "nsk_ctx_create": (C.c_int, [C.c_int, c_void_pp]),
    "nsk_ctx_destroy": (C.c_int, [C.c_void_p]),
    "nsk_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsk_ctx_get_stream": (C.c_void_p, [C.c_void_p]),
    "nsk_ctx_sync": (C.c_int, [C.c_void_p]),
    "nsk_ctx_launch_count": (C.c_uint64, [C.c_void_p]),
    "nsk_ctx_device_info": (C.c_int, [C.c_void_p, c_int_p, c_int64_p, c_int_p, c_int64_p]),
    "nsk_ctx_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "nsk_event_create": (C.c_int, [C.c_void_p, c_void_pp]),
    "nsk_event_destroy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsk_event_record": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsk_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_float)]),
    "nsk_malloc": (C.c_int, [C.c_void_p, C.c_size_t, c_void_pp]),
    "nsk_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsk_host_alloc": (C.c_int, [C.c_void_p, C.c_size_t, c_void_pp]),
    "nsk_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "nsk_memcpy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]),
    "nsk_memset0": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "nsk_flush_l2": (C.c_int, [C.c_void_p]),
    "nsk_ctx_query": (C.c_int, [C.c_void_p, C.c_char_p, c_int64_p]),
    "nsk_coo2csr": (C.c_int64, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsk_coo2bcsr4": (C.c_int64, [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsk_mtx_read": (C.c_int, [C.c_char_p, c_int_p, c_int64_p, C.POINTER(C.POINTER(C.c_int)), C.POINTER(C.POINTER(C.c_int)),
                               C.POINTER(C.POINTER(C.c_double))]),
    "nsk_mtx_free": (None, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "nsk_rcm": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsk_csr_permute": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsk_csr_bandwidth": (C.c_int64, [C.c_int, C.c_void_p, C.c_void_p]),
    "nsk_pack_host_create": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, c_void_pp]),
    "nsk_pack_host_why": (C.c_char_p, [C.c_void_p]),
    "nsk_pack_host_bytes": (C.c_int64, [C.c_void_p]),
    "nsk_pack_host_expand": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_int_p, c_int_p]),
    "nsk_pack_host_simulate": (C.c_longlong, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                              C.c_uint, C.POINTER(C.c_longlong), c_int_p, C.c_int]),
    "nsk_pack_host_destroy": (None, [C.c_void_p]),
    "nsk_sell_host_create": (C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, c_void_pp]),
    "nsk_sell_host_why": (C.c_char_p, [C.c_void_p]),
    "nsk_sell_host_global_pattern": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsk_sell_host_stats": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p, c_int64_p]),
    "nsk_sell_host_expand": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsk_sell_host_simulate": (C.c_longlong, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                              C.c_void_p, C.c_uint, C.POINTER(C.c_longlong), c_int_p, C.c_int]),
    "nsk_sell_host_destroy": (None, [C.c_void_p]),
    "nsk_csr_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                 c_void_pp]),
    "nsk_csr_destroy": (C.c_int, [C.c_void_p]),
    "nsk_csr_shape": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int64_p]),
    "nsk_csr_spmv_bytes": (C.c_int64, [C.c_void_p]),
    "nsk_csr_mpk_bytes": (C.c_int64, [C.c_void_p, C.c_int]),
    "nsk_csr_packed_bytes": (C.c_int64, [C.c_void_p]),
    "nsk_csr_tile_bytes": (C.c_int64, [C.c_void_p]),
    "nsk_spmv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "nsk_mpk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, c_void_pp, C.c_int, C.c_int]),
    "nsk_mpk_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_void_pp, c_void_pp, C.c_int, C.c_int]),
    "nsk_bcsr4_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, c_void_pp]),
    "nsk_bcsr4_destroy": (C.c_int, [C.c_void_p]),
    "nsk_spmv_bcsr4": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "nsk_bcsr4_mpk": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "nsk_spmm_bcsr4": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int]),
    "nsk_krylov_basis_bcsr4": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]),
    "nsk_dot": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, c_double_p, C.c_int]),
    "nsk_norm2": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, c_double_p, C.c_int]),
    "nsk_rel_error": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, c_double_p, C.c_int]),
    "nsk_axpy": (C.c_int, [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_int]),
    "nsk_orthogonalize": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double, c_double_p, C.c_int]),
    "nsk_orthonormalize_against_basis": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_void_pp, C.c_void_p, c_double_p, C.c_int]),
    "nsk_gram": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, c_void_pp, c_double_p, C.c_int]),
    "nsk_cg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, c_int_p, c_double_p,
                         C.c_int]),
    "nsk_comm_unique_id": (C.c_int, [C.c_void_p]),
    "nsk_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "nsk_comm_destroy": (C.c_int, [C.c_void_p]),
    "nsk_comm_allreduce_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "nsk_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_int, c_void_pp]),
    "nsk_plan_destroy": (C.c_int, [C.c_void_p]),
    "nsk_plan_frontier": (C.c_int, [C.c_void_p, c_int_p, c_void_pp]),
    "nsk_plan_add_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsk_plan_finalize": (C.c_int, [C.c_void_p]),
    "nsk_plan_sizes": (C.c_int, [C.c_void_p, c_int_p, c_int_p, c_int_p, c_int64_p, C.c_void_p, C.c_void_p]),
    "nsk_plan_ghosts": (C.c_int, [C.c_void_p, c_void_pp]),
    "nsk_plan_local_csr": (C.c_int, [C.c_void_p, c_void_pp, c_void_pp, c_void_pp]),
    "nsk_plan_requests": (C.c_int, [C.c_void_p, C.c_int, c_int_p, c_void_pp, C.c_void_p]),
    "nsk_plan_add_send": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "nsk_csr_create_dist": (C.c_int, [C.c_void_p, C.c_void_p, c_void_pp]),
    "nsk_csr_owned_rows": (C.c_int, [C.c_void_p]),
    "nsk_halo_exchange": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "nsk_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsk_ipc_import": (C.c_int, [C.c_void_p, C.c_void_p, c_void_pp]),
    "nsk_dist_push_flags": (C.c_int, [C.c_void_p, c_void_pp]),
    "nsk_dist_recv_layout": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "nsk_dist_peer_count": (C.c_int, [C.c_void_p]),
    "nsk_dist_peer_rank": (C.c_int, [C.c_void_p, C.c_int]),
    "nsk_dist_push_peer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "nsk_dist_vector_register": (C.c_int, [C.c_void_p, C.c_void_p, c_void_pp]),
}

_lib = None


class NskError(RuntimeError):
    def __init__(self, status: int, detail: str):
        super().__init__(f"nsk status {status}: {detail}")
        self.status = status
        self.detail = detail


def load() -> C.CDLL:
    """Loads libnsk.so (raises if it has not been built: run `python -m navierstokes_b200.build`)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} not found: build it with `python -m navierstokes_b200.build` "
                "(there is no CPU fallback for this path)")
        lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int, ctx=None) -> int:
    if status < 0 and status != -7:
        lib = load()
        detail = lib.nsk_last_error(ctx) or b""
        raise NskError(status, f"{lib.nsk_strerror(status).decode()}: {detail.decode(errors='replace')}")
    return status
