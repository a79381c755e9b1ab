"""ctypes mirrors of the C ABI in csrc/ (structs in csrc/include/afb_round.cuh, afb_common.cuh).

Keep field order/types in lock-step with the headers; ``tests/test_abi.py`` checks the sizes
against ``afb_abi_sizeof_*`` exported by the library.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

vp, i32, i64, f32, u32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32, C.c_uint64


class RoundParams(C.Structure):
    _fields_ = [
        ("X", vp), ("indptr", vp), ("indices", vp), ("values", vp),
        ("y", vp), ("gidx", vp), ("hist", vp), ("hist_prev", vp), ("hist_tag", vp),
        ("row_base", i64), ("ld", i64),
        ("w_server", vp), ("abar_server", vp), ("sctl", vp), ("status", vp), ("ring", vp),
        ("w_local", vp), ("abar_local", vp), ("g_local", vp), ("wctl", vp), ("peers", vp),
        ("n_rows", i32), ("d", i32), ("algo", i32), ("ring_size", i32),
        ("num_workers", i32), ("worker_id", i32), ("thr24", i32), ("gate_threshold", i32),
        ("seed", i64), ("taw", i64), ("max_iters", i64),
        ("gamma", f32), ("inv_par_recs", f32), ("inv_n", f32), ("delay_coeff", f32),
        ("delay_mode", i32), ("warmup_updates", i32), ("seed_mode", i32), ("loss_id", i32),
        ("round_host", i32), ("pdl", i32),
        ("reg_param", f32), ("reg_type", i32), ("ssp_bound", i32), ("use_mask", i32),
        ("desc", vp), ("desc_slots", i32), ("pull_ahead_chunks", i32), ("spin_limit_ns", u64),
        ("tma_mask", i32), ("g_copies", i32),
    ]


REG_TYPES = {None: 0, "none": 0, "l2": 1, "squared_l2": 1, "l1": 2}

# per-round descriptor of the end-to-end path (afb::RoundDesc): {gamma, thr24, delay_coeff, seq}
ROUND_DESC_DTYPE = np.dtype([("gamma", "<f4"), ("thr24", "<i4"), ("delay_coeff", "<f4"), ("seq", "<u4")], align=False)
assert ROUND_DESC_DTYPE.itemsize == 16

# double-buffered per-round state of the pipelined persistent rounds (afb::PipeBuf x 2)
PIPE_BUF_DTYPE = np.dtype([
    ("chunk_ctr", "<u4"), ("done_ctas", "<u4"), ("pull_ticket", "<u4"), ("pull_flag", "<u4"),
    ("proceed", "<i4"), ("batch_count", "<u4"), ("loss_sum", "<f4"), ("tk_flag", "<u4"),
    ("ts", "<u8"), ("k_read", "<u8"), ("t_start_ns", "<u8"), ("tk_staleness", "<i8"), ("tk_k_after", "<i8"),
    ("tk_accept", "<i4"), ("tk_step", "<f4"), ("tk_clock_old", "<u4"), ("pad", "<u4"), ("pad16", "<u8"),
], align=False)
assert PIPE_BUF_DTYPE.itemsize == 96      # alignas(16): 88 bytes of fields + tail padding


class SyncPeers(C.Structure):
    _fields_ = [("gstage", vp * 16), ("flags", vp * 16), ("mc_gstage", vp), ("mc_flags", vp)]


# numpy views of the device-side control blocks (for host reads after a D2H copy)
WORKER_CTL_DTYPE = np.dtype([
    ("chunk_ctr", "<u4"), ("done_ctas", "<u4"), ("pull_ticket", "<u4"), ("pull_flag", "<u4"),
    ("round", "<u4"), ("proceed", "<i4"), ("batch_count", "<u4"), ("loss_sum", "<f4"),
    ("ts", "<u8"), ("k_read", "<u8"), ("seed_key", "<u8"), ("t_start_ns", "<u8"),
    ("last_accepted", "<i4"), ("last_batch", "<u4"), ("last_loss", "<f4"), ("last_step", "<f4"),
    ("last_staleness", "<i8"), ("last_k_after", "<i8"), ("last_task_ns", "<u8"), ("rounds_done", "<u8"),
    ("wait_ns", "<u8"), ("last_finish_ns", "<u8"), ("avg_task_ms", "<f4"), ("pad_f", "<f4"), ("pad_u", "<u8"),
    ("dbg_ns", "<u8", (8,)),
    ("tk_flag", "<u4"), ("tk_accept", "<i4"), ("tk_step", "<f4"), ("tk_clock_old", "<u4"),
    ("tk_staleness", "<i8"), ("tk_k_after", "<i8"),
    ("abort", "<i4"), ("skipped_rounds", "<u4"), ("wave_seen", "<u4"), ("pad_tail", "<u4"),
], align=False)
assert WORKER_CTL_DTYPE.itemsize == 256

SERVER_CTL_DTYPE = np.dtype({
    "names": ["clock_k", "avail_count", "stop", "cul_time_ns", "cul_count", "wave", "avg_delay_ns", "gate_dyn", "num_alive",
              "enable_mask"],
    "formats": ["<u8", "<i4", "<i4", "<u8", "<u4", "<u4", "<u8", "<i4", "<i4", ("<u8", (4,))],
    "offsets": [0, 8, 12, 16, 24, 28, 32, 40, 44, 48], "itemsize": 80})

WORKER_SLOT_DTYPE = np.dtype([
    ("available", "<i4"), ("alive", "<i4"), ("staleness", "<i8"), ("num_tasks", "<i8"),
    ("avg_task_ms", "<f4"), ("last_task_ms", "<f4"), ("heartbeat_ns", "<u8"), ("busy_since_ns", "<u8"),
    ("wait_ns", "<u8"), ("last_finish_ns", "<u8"),
], align=False)
assert WORKER_SLOT_DTYPE.itemsize == 64

RESULT_REC_DTYPE = np.dtype({
    "names": ["worker", "accepted", "ts", "staleness", "k_after", "batch", "loss_sum", "step",
              "t_start_ns", "t_end_ns", "seq"],
    "formats": ["<i4", "<i4", "<i8", "<i8", "<i8", "<i8", "<f4", "<f4", "<u8", "<u8", "<u8"],
    "offsets": [0, 4, 8, 16, 24, 32, 40, 44, 48, 56, 64], "itemsize": 80})


def declare(lib: C.CDLL) -> None:
    def sig(name, restype, argtypes):
        if not hasattr(lib, name):
            return
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes

    P = C.POINTER
    sig("afb_dense_stream_launch", C.c_int, [P(RoundParams), C.c_int, C.c_int, C.c_int, C.c_int, vp])
    sig("afb_dense_stream_pipelined_launch", C.c_int, [P(RoundParams), C.c_int, C.c_int, vp, vp, vp, C.c_int, vp])
    sig("afb_dense_stream_max_ld", C.c_int, [])
    sig("afb_dense_stream_ctas_per_sm", C.c_int, [i64, C.c_int])
    sig("afb_dense_stream_depth", C.c_int, [i64, C.c_int])
    sig("afb_pipe_state_bytes", C.c_int, [])
    sig("afb_dense_round_launch", C.c_int, [P(RoundParams), C.c_int, vp])
    sig("afb_dense_round_max_ld", C.c_int, [])
    sig("afb_dense_round_ctas_per_sm", C.c_int, [i64])
    sig("afb_csr_round_launch", C.c_int, [P(RoundParams), C.c_int, vp])
    sig("afb_dense_round_pipe_launch", C.c_int, [P(RoundParams), C.c_int, vp])
    sig("afb_dense_round_pipe_max_ld", C.c_int, [])
    sig("afb_dense_round_ring_launch", C.c_int, [P(RoundParams), C.c_int, vp])
    sig("afb_dense_round_ring_max_ld", C.c_int, [])
    sig("afb_sleep", C.c_int, [i64, vp])
    sig("afb_e2e_create", vp, [])
    sig("afb_e2e_destroy", None, [vp])
    sig("afb_e2e_pre", C.c_int, [vp, C.c_int, vp, vp, i64, vp])
    sig("afb_e2e_post", C.c_int, [vp, C.c_int, vp, vp, vp, i64, vp])
    sig("afb_e2e_wait", C.c_int, [vp, C.c_int])
    sig("afb_e2e_sync", C.c_int, [vp])
    sig("afb_saga_rollback_launch", C.c_int, [vp, vp, vp, C.c_int, vp, vp])
    sig("afb_axpy", C.c_int, [f32, vp, vp, i64, vp])
    sig("afb_scal", C.c_int, [f32, vp, i64, vp])
    sig("afb_dot", C.c_int, [vp, vp, i64, vp, vp])
    sig("afb_updater", C.c_int, [C.c_int, vp, vp, vp, i64, f32, f32, vp, vp])
    sig("afb_sample_rows_v2", C.c_int, [vp, i64, i64, i64, C.c_int, vp, vp, vp, vp])
    sig("afb_eval_mse_dense", C.c_int, [vp, vp, i64, i64, C.c_int, vp, C.c_int, C.c_int, vp, vp])
    sig("afb_eval_mse_csr", C.c_int, [vp, vp, vp, vp, i64, C.c_int, vp, C.c_int, C.c_int, vp, vp])
    sig("afb_eval_mse_tc", C.c_int, [vp, vp, i64, i64, C.c_int, vp, C.c_int, vp, vp, vp])
    sig("afb_eval_loss_tc", C.c_int, [vp, vp, i64, i64, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp])
    sig("afb_csr_stream_launch", C.c_int, [P(RoundParams), C.c_int, C.c_int, C.c_int, vp, vp])
    sig("afb_grid_bar_bytes", C.c_int, [])
    sig("afb_abi_sizeof", C.c_int, [C.c_int])
    sig("afb_ipc_get_handle", C.c_int, [vp, vp])
    sig("afb_ipc_open_handle", C.c_int, [vp, P(vp)])
    sig("afb_ipc_close_handle", C.c_int, [vp])
    sig("afb_enable_peer_access", C.c_int, [C.c_int, C.c_int])
    sig("afb_p2p_pingpong", C.c_int, [vp, vp, C.c_int, C.c_int, vp, vp])
    sig("afb_push_apply_launch", C.c_int, [P(RoundParams), vp, C.c_int, vp])
    sig("afb_gate_launch", C.c_int, [P(RoundParams), vp])
    sig("afb_pull_w_launch", C.c_int, [P(RoundParams), C.c_int, vp])
    sig("afb_broadcast_mc", C.c_int, [vp, vp, i64, vp, vp, vp])
    sig("afb_broadcast_p2p", C.c_int, [vp, vp, vp, C.c_int, i64, vp, vp])
    sig("afb_l2_flush", C.c_int, [vp, i64, vp])
    sig("afb_peer_copy", C.c_int, [vp, vp, i64, vp])
