from .vectors import DenseVector, LabeledPoint, SparseVector, Vector, Vectors
from .matrices import DenseMatrix, Matrices, Matrix, SparseMatrix
from . import blas_util, converters, matrices
from .blas_util import axpy_op, dot_op, scal_op
from .matrices import gemm, gemv, spr, syr

def __getattr__(name):                 # RowMatrix needs data/ (which imports linalg): resolve lazily to avoid the cycle
    if name == "RowMatrix":
        from .distributed import RowMatrix
        return RowMatrix
    raise AttributeError(name)


__all__ = ["RowMatrix", "DenseVector", "LabeledPoint", "SparseVector", "Vector", "Vectors", "DenseMatrix", "SparseMatrix", "Matrix", "Matrices",
           "blas_util", "converters", "matrices", "axpy_op", "dot_op", "scal_op", "gemv", "gemm", "syr", "spr"]
