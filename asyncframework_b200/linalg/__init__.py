from .vectors import DenseVector, LabeledPoint, SparseVector, Vector, Vectors
from .matrices import DenseMatrix, Matrices, Matrix, SparseMatrix
from . import blas_util, converters, matrices
from .blas_util import axpy_op, dot_op, scal_op
from .matrices import gemm, gemv, spr, syr

__all__ = ["DenseVector", "LabeledPoint", "SparseVector", "Vector", "Vectors", "DenseMatrix", "SparseMatrix", "Matrix", "Matrices",
           "blas_util", "converters", "matrices", "axpy_op", "dot_op", "scal_op", "gemv", "gemm", "syr", "spr"]
