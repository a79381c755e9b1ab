from .vectors import DenseVector, LabeledPoint, SparseVector, Vector, Vectors
from . import blas_util, converters
from .blas_util import axpy_op, dot_op, scal_op

__all__ = ["DenseVector", "LabeledPoint", "SparseVector", "Vector", "Vectors",
           "blas_util", "converters", "axpy_op", "dot_op", "scal_op"]
