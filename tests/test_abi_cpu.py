"""ABI lock-step between the C structs in csrc/include and their ctypes/numpy mirrors — checked on the CPU
(libafb_kernels.so links cudart statically and resolves the driver lazily, so it loads without a GPU)."""
import ctypes as C

import pytest


def test_struct_sizes_match_native_library():
    from asyncframework_b200 import _native
    from asyncframework_b200.ops import _signatures as S
    if not _native.kernels_path().exists():
        pytest.skip("libafb_kernels.so not built")
    lib = _native.kernels()
    assert lib.afb_abi_sizeof(0) == C.sizeof(S.RoundParams)
    assert lib.afb_abi_sizeof(1) == S.WORKER_CTL_DTYPE.itemsize
    assert lib.afb_abi_sizeof(2) == S.SERVER_CTL_DTYPE.itemsize
    assert lib.afb_abi_sizeof(3) == S.WORKER_SLOT_DTYPE.itemsize
    assert lib.afb_abi_sizeof(4) == S.RESULT_REC_DTYPE.itemsize
    assert lib.afb_abi_sizeof(5) == C.sizeof(S.SyncPeers)


def test_every_declared_entry_point_exists():
    """A symbol renamed in csrc/ but not in _signatures.py (or vice versa) must fail here, not on the GPU box."""
    from asyncframework_b200 import _native
    if not _native.kernels_path().exists():
        pytest.skip("libafb_kernels.so not built")
    lib = _native.kernels()
    for name in ("afb_dense_round_launch", "afb_dense_round_max_ld", "afb_dense_round_ctas_per_sm", "afb_csr_round_launch",
                 "afb_dense_round_ring_launch", "afb_dense_round_pipe_launch", "afb_saga_rollback_launch", "afb_axpy", "afb_scal",
                 "afb_dot", "afb_updater", "afb_sample_rows_v2", "afb_eval_mse_dense", "afb_eval_mse_csr", "afb_eval_mse_tc",
                 "afb_ipc_get_handle", "afb_ipc_open_handle", "afb_enable_peer_access", "afb_p2p_pingpong", "afb_peer_copy",
                 "afb_push_apply_launch", "afb_gate_launch", "afb_pull_w_launch", "afb_broadcast_mc", "afb_broadcast_p2p",
                 "afb_l2_flush", "afb_sleep", "afb_abi_sizeof"):
        fn = getattr(lib, name)
        assert fn.argtypes is not None, f"{name} has no declared signature"
    assert lib.afb_dense_round_max_ld() == 2048 and lib.afb_dense_round_ctas_per_sm(784) == 2 and lib.afb_dense_round_ctas_per_sm(2000) == 1


def test_stream_kernels_are_in_the_default_library():
    """The persistent round kernels (dense TMA-ring stream kernel, one-kernel CSR round) and the tensor-core evaluation with a
    loss epilogue ship in the default library: no opt-in side library exists any more."""
    from asyncframework_b200 import _native
    from asyncframework_b200.ops import _signatures as S
    if not _native.kernels_path().exists():
        pytest.skip("libafb_kernels.so not built")
    lib = _native.kernels()
    for name in ("afb_dense_stream_launch", "afb_dense_stream_pipelined_launch", "afb_csr_stream_launch", "afb_eval_loss_tc",
                 "afb_pipe_state_bytes", "afb_grid_bar_bytes"):
        assert getattr(lib, name).argtypes is not None, name
    assert lib.afb_dense_stream_max_ld() == 2048
    assert lib.afb_dense_stream_ctas_per_sm(784, 0) == 2 and lib.afb_dense_stream_ctas_per_sm(2000, 0) == 1
    assert lib.afb_dense_stream_depth(784, 0) == 4 and lib.afb_dense_stream_depth(784, 1) == 8 and lib.afb_dense_stream_depth(2000, 0) == 3
    assert lib.afb_abi_sizeof(6) == S.ROUND_DESC_DTYPE.itemsize and lib.afb_abi_sizeof(7) == S.PIPE_BUF_DTYPE.itemsize
    assert lib.afb_pipe_state_bytes() == 2 * S.PIPE_BUF_DTYPE.itemsize
    assert not (_native.LIBDIR / "libafb_experimental.so").exists()
