"""BLASSuite-style checks (mllib test BLASSuite.scala:26-299): copy/scal/axpy/dot, dense + sparse."""
import pytest
import numpy as np
import torch

from asyncframework_b200.linalg import DenseVector, LabeledPoint, SparseVector, Vectors, blas_util
from asyncframework_b200.linalg.blas_util import axpy, axpy_op, copy, dot, dot_op, scal, scal_op


def test_factories_and_equality():
    sv = Vectors.sparse(4, [0, 2], [1.0, -2.0])
    dv = Vectors.dense(1.0, 0.0, -2.0, 0.0)
    assert sv == dv and sv.to_dense() == dv and dv.to_sparse().indices.tolist() == [0, 2]
    assert Vectors.sparse(3, [(2, 5.0), (0, 1.0)]).indices.tolist() == [0, 2]
    assert Vectors.zeros(3).values.tolist() == [0, 0, 0]
    assert sv.num_nonzeros == 2 and len(sv) == 4
    with pytest.raises(ValueError):
        SparseVector(3, [2, 1], [1.0, 1.0])
    with pytest.raises(ValueError):
        SparseVector(3, [3], [1.0])
    lp = LabeledPoint(1.0, dv)
    assert tuple(lp)[0] == 1.0


def test_axpy_dense_and_sparse():
    y = Vectors.dense(1.0, 2.0, 3.0)
    axpy(2.0, Vectors.dense(1.0, 1.0, 1.0), y)
    assert y.values.tolist() == [3.0, 4.0, 5.0]
    axpy(-1.0, Vectors.sparse(3, [1], [4.0]), y)
    assert y.values.tolist() == [3.0, 0.0, 5.0]
    with pytest.raises(ValueError):
        axpy(1.0, Vectors.dense(1.0), y)
    with pytest.raises(TypeError):
        axpy(1.0, y, Vectors.sparse(3, [0], [1.0]))


def test_dot_all_combinations():
    sx = Vectors.sparse(4, [0, 2, 3], [1.0, 2.0, 3.0])
    sy = Vectors.sparse(4, [1, 2, 3], [5.0, 7.0, 11.0])
    dx, dy = sx.to_dense(), sy.to_dense()
    expect = 2 * 7 + 3 * 11
    assert dot(dx, dy) == expect and dot(sx, dy) == expect and dot(dx, sy) == expect and dot(sx, sy) == expect
    assert dot(Vectors.sparse(4, [0], [1.0]), Vectors.sparse(4, [1], [1.0])) == 0.0


def test_scal_copy_and_chaining():
    x = Vectors.dense(1.0, -2.0)
    scal(3.0, x)
    assert x.values.tolist() == [3.0, -6.0]
    s = Vectors.sparse(3, [1], [2.0])
    scal(0.5, s)
    assert s.values.tolist() == [1.0]
    y = Vectors.zeros(2)
    copy(x, y)
    assert y == x
    # BLASUtil chaining: axpyOp returns the mutated y, scalOp returns x (BLASUtil.scala:6-18)
    g = Vectors.zeros(2)
    out = axpy_op(2.0, scal_op(0.5, Vectors.dense(2.0, 4.0)), g)
    assert out is g and g.values.tolist() == [2.0, 4.0] and dot_op(g, g) == 20.0


# ---- level 2/3 + matrices (BLASSuite.scala: gemm / gemv / syr / spr, dense + sparse; MatricesSuite) -----------------------
def test_matrices_storage_and_conversions():
    from asyncframework_b200.linalg import DenseMatrix, Matrices, SparseMatrix
    # 3x2 column-major [1 4; 2 5; 3 6]
    A = Matrices.dense(3, 2, [1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    assert A[0, 1] == 4.0 and A.to_tensor().tolist() == [[1, 4], [2, 5], [3, 6]]
    assert A.transpose.to_tensor().tolist() == [[1, 2, 3], [4, 5, 6]] and A.transpose.transpose[2, 1] == 6.0
    with pytest.raises(ValueError):
        DenseMatrix(2, 2, [1.0, 2.0, 3.0])
    # the MatricesSuite sparse example: 3x2 with col 0 = {0:1, 1:2}, col 1 = {1:4, 2:5}... as CSC
    S = Matrices.sparse(3, 2, [0, 2, 4], [0, 1, 1, 2], [1.0, 2.0, 4.0, 5.0])
    assert S.to_tensor().tolist() == [[1, 0], [2, 4], [0, 5]] and S.num_nonzeros == 4
    assert S.transpose.to_tensor().tolist() == [[1, 2, 0], [0, 4, 5]]
    back = S.to_dense().to_sparse()
    assert back.col_ptrs.tolist() == [0, 2, 4] and back.row_indices.tolist() == [0, 1, 1, 2]
    with pytest.raises(ValueError):
        SparseMatrix(3, 2, [0, 2], [0, 1], [1.0, 2.0])
    assert Matrices.eye(3).to_tensor().equal(torch.eye(3, dtype=torch.float64)) and Matrices.speye(3).to_tensor().equal(torch.eye(3, dtype=torch.float64))
    assert Matrices.horzcat([A, A]).num_cols == 4 and Matrices.vertcat([A, A]).num_rows == 6
    assert Matrices.diag(Vectors.dense(1.0, 2.0)).to_tensor().tolist() == [[1, 0], [0, 2]]
    assert SparseMatrix.from_coo(2, 2, [(0, 0, 1.0), (0, 0, 2.0), (1, 1, 5.0)]).to_tensor().tolist() == [[3, 0], [0, 5]]


def test_gemv_gemm_dense_sparse_and_transposed_match_numpy():
    from asyncframework_b200.linalg import DenseMatrix, SparseMatrix, gemm, gemv
    rng = np.random.default_rng(0)
    M = rng.standard_normal((5, 4)) * (rng.random((5, 4)) < 0.5)
    B = rng.standard_normal((4, 3))
    xd = rng.standard_normal(4)
    dA, sA = DenseMatrix.from_tensor(torch.from_numpy(M)), SparseMatrix.from_dense(torch.from_numpy(M))
    dB = DenseMatrix.from_tensor(torch.from_numpy(B))
    xs = Vectors.sparse(4, [1, 3], [xd[1], xd[3]])
    xs_dense = np.array([0, xd[1], 0, xd[3]])
    for A in (dA, sA):
        for x, xn in ((Vectors.dense(xd), xd), (xs, xs_dense)):
            y0 = rng.standard_normal(5)
            y = Vectors.dense(y0.copy())
            gemv(2.0, A, x, 0.5, y)
            assert np.allclose(y.to_array(), 2.0 * M @ xn + 0.5 * y0)
        C0 = rng.standard_normal((5, 3))
        C = DenseMatrix.from_tensor(torch.from_numpy(C0.copy()))
        gemm(1.5, A, dB, -1.0, C)
        assert np.allclose(C.to_array(), 1.5 * M @ B - C0)
        # transposed operands (A^T is 4x5): (A^T)^T B through the views, no copies of the data
        Ct = DenseMatrix.zeros(5, 3)
        gemm(1.0, A.transpose.transpose, dB, 0.0, Ct)
        assert np.allclose(Ct.to_array(), M @ B)
        yt = Vectors.dense(np.zeros(4))
        gemv(1.0, A.transpose, Vectors.dense(np.ones(5)), 0.0, yt)
        assert np.allclose(yt.to_array(), M.T @ np.ones(5))
        assert np.allclose(A.multiply(dB).to_array(), M @ B) and np.allclose(A.multiply(Vectors.dense(xd)).to_array(), M @ xd)
    y = Vectors.dense(np.ones(5))
    gemv(0.0, dA, Vectors.dense(xd), 1.0, y)                        # alpha = 0, beta = 1: untouched
    assert y.to_array().tolist() == [1.0] * 5
    with pytest.raises(ValueError):
        gemv(1.0, dA, Vectors.dense(np.ones(5)), 0.0, Vectors.dense(np.ones(5)))
    with pytest.raises(ValueError):
        gemm(1.0, dA, dB, 0.0, DenseMatrix.zeros(3, 5).transpose)


def test_syr_and_spr_rank_one_updates():
    from asyncframework_b200.linalg import DenseMatrix, spr, syr
    from asyncframework_b200.linalg.matrices import unpack_upper
    # BLASSuite "syr": A = [[0,1.2,2.2,3.1],...] symmetric, x = (0, 2.7, 3.5, 2.1), alpha = 0.15
    A0 = np.array([[0.0, 1.2, 2.2, 3.1], [1.2, 3.2, 5.3, 4.6], [2.2, 5.3, 1.8, 3.0], [3.1, 4.6, 3.0, 0.8]])
    x = np.array([0.0, 2.7, 3.5, 2.1])
    for xv in (Vectors.dense(x), Vectors.sparse(4, [1, 2, 3], x[1:])):
        A = DenseMatrix.from_tensor(torch.from_numpy(A0.copy()))
        syr(0.15, xv, A)
        assert np.allclose(A.to_array(), A0 + 0.15 * np.outer(x, x))
    with pytest.raises(ValueError):
        syr(1.0, Vectors.dense(x), DenseMatrix.zeros(4, 3))
    # "spr": packed upper triangle, column-major
    for xv in (Vectors.dense(x), Vectors.sparse(4, [1, 2, 3], x[1:])):
        U = torch.tensor([A0[i, j] for j in range(4) for i in range(j + 1)])
        spr(0.1, xv, U)
        assert np.allclose(unpack_upper(U, 4).numpy(), A0 + 0.1 * np.outer(x, x))
    with pytest.raises(ValueError):
        spr(1.0, Vectors.dense(x), torch.zeros(9, dtype=torch.float64))


def test_row_matrix_gramian_covariance_pca_svd_multiply_match_numpy():
    """RowMatrixSuite at the behaviour level, dense and CSR shards."""
    from asyncframework_b200.core.scheduler import Scheduler
    from asyncframework_b200.data import ShardedDataset
    from asyncframework_b200.linalg import DenseMatrix, RowMatrix
    rng = np.random.default_rng(6)
    A = rng.standard_normal((400, 5)) @ np.diag([5.0, 2.0, 1.0, 0.5, 0.1]) + np.array([1.0, -2.0, 0.0, 3.0, 0.5])
    with Scheduler(3) as s:
        dense = ShardedDataset.from_arrays(s, A, np.zeros(400), 3)
        order = np.concatenate([np.arange(p, 400, 3) for p in range(3)])
        for ds in (dense, ShardedDataset.from_shards(s, [sh.to_csr() for sh in dense.shards()])):
            M = RowMatrix(ds)
            assert (M.num_rows(), M.num_cols()) == (400, 5)
            assert np.allclose(M.compute_gramian_matrix().to_array(), A.T @ A)
            assert np.allclose(M.compute_covariance().to_array(), np.cov(A, rowvar=False))
            pcs, var = M.compute_principal_components_and_explained_variance(2)
            w, v = np.linalg.eigh(np.cov(A, rowvar=False))
            top = v[:, np.argsort(w)[::-1][:2]]
            assert np.allclose(np.abs(pcs.to_array().T @ top), np.eye(2), atol=1e-8)          # same subspace directions (sign-free)
            assert np.allclose(var.to_array(), np.sort(w)[::-1][:2] / w.sum())
            U, sv, V = M.compute_svd(3, compute_u=True)
            assert np.allclose(sv.to_array(), np.linalg.svd(A, compute_uv=False)[:3])
            recon = U.to_tensor().numpy() @ np.diag(sv.to_array()) @ V.to_array().T
            best3 = (lambda u, s_, vt: (u[:, :3] * s_[:3]) @ vt[:3])(*np.linalg.svd(A[order], full_matrices=False))
            assert np.allclose(recon, best3, atol=1e-8)                                      # the rank-3 truncation of A
            B = rng.standard_normal((5, 2))
            assert np.allclose(M.multiply(DenseMatrix.from_tensor(torch.from_numpy(B))).to_tensor().numpy(), A[order] @ B)
            st = M.compute_column_summary_statistics()
            assert np.allclose(st.mean.numpy(), A.mean(0))
        with pytest.raises(ValueError):
            RowMatrix(dense).compute_svd(6)
        with pytest.raises(ValueError):
            RowMatrix(dense).multiply(DenseMatrix.zeros(4, 2))


def test_vector_and_labeled_point_text_forms_round_trip():
    """Vectors.parse / LabeledPoint.parse (NumericParser grammar) read back what str() — Spark's saveAsTextFile form — writes."""
    import pytest
    from asyncframework_b200.linalg.vectors import LabeledPoint, Vectors
    dv = Vectors.dense(1.5, -2.0, 0.0, 3e-7)
    sv = Vectors.sparse(6, [1, 4], [2.5, -1.0])
    assert str(dv) == "[1.5,-2.0,0.0,3e-07]" and str(sv) == "(6,[1,4],[2.5,-1.0])"
    assert Vectors.parse(str(dv)) == dv and Vectors.parse(str(sv)) == sv and Vectors.parse(" [ 1.0 , 2.0 ] ") == Vectors.dense(1.0, 2.0)
    assert Vectors.parse("[]").size == 0 and Vectors.parse("(3,[],[])") == Vectors.sparse(3, [], [])
    for lp in (LabeledPoint(1.0, dv), LabeledPoint(-3.5, sv)):
        back = LabeledPoint.parse(str(lp))
        assert back.label == lp.label and back.features == lp.features
    old = LabeledPoint.parse("1.0,0.5 2.5 -1")                     # the pre-1.0 dense line format
    assert old.label == 1.0 and old.features == Vectors.dense(0.5, 2.5, -1.0)
    for bad in ("[1.0,2.0", "(1.0,[1.0,x])", "[1.0](", "(4,[0],[1.0],[2.0])", "(1.0)"):
        with pytest.raises(ValueError):
            (LabeledPoint.parse if bad.startswith("(1.0") else Vectors.parse)(bad)
