"""BLASSuite-style checks (mllib test BLASSuite.scala:26-299): copy/scal/axpy/dot, dense + sparse."""
import pytest
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
