"""Generalised linear models trained with mini-batch SGD (the MLlib surface ASYNC touches).

Parity: ``GeneralizedLinearAlgorithm.run`` (mllib/regression/GeneralizedLinearAlgorithm.scala:
248-360: validators, optional intercept as a trailing 1.0 feature :291-332, optional feature
scaling :283-304), ``LinearRegressionWithSGD`` (LeastSquaresGradient + SimpleUpdater,
regression/LinearRegression.scala:92-107) **with the ASYNC weight history ``.h``** (:130,178-184),
``LogisticRegressionWithSGD`` (LogisticGradient + SquaredL2Updater, binary-label validator,
classification/LogisticRegression.scala:217-243), ``SVMWithSGD``, ``RidgeRegressionWithSGD``,
``LassoWithSGD``; models with ``predict`` / threshold handling (:88-132) and text save/load
(util/modelSaveLoad.scala).  Python-facing ``train(...)`` signatures follow PySpark
(python/pyspark/mllib/regression.py:229, classification.py:269).
"""
from __future__ import annotations

import json
import os
from typing import List, Optional, Tuple

import torch

from ..data.batch import Shard
from ..data.dataset import ShardedDataset
from ..linalg.vectors import DenseVector, LabeledPoint, Vector
from ..optim.gradient import HingeGradient, LeastSquaresGradient, LogisticGradient
from ..optim.gradient_descent import GradientDescent
from ..optim.updater import L1Updater, SimpleUpdater, SquaredL2Updater, make_updater


# ------------------------------------------------------------------------------------------ models
class GeneralizedLinearModel:
    kind = "glm"

    def __init__(self, weights: torch.Tensor, intercept: float = 0.0):
        self.weights = weights.detach().clone()
        self.intercept = float(intercept)

    def _margin(self, x):
        if isinstance(x, Shard):
            return x.matvec(self.weights.to(x.device)) + self.intercept
        if isinstance(x, Vector):
            from ..linalg import blas_util
            return blas_util.dot(x, DenseVector(self.weights, dtype=self.weights.dtype)) + self.intercept
        xt = torch.as_tensor(x, dtype=self.weights.dtype)
        return xt @ self.weights.to(xt.device) + self.intercept

    def predict_point(self, margin):
        return margin

    def predict(self, x):
        if isinstance(x, ShardedDataset):
            return x.map_batch(lambda s: self.predict_point(self._margin(s)))
        return self.predict_point(self._margin(x))

    # --- persistence: MLlib-compatible *content* (class, weights, intercept[, threshold]) as text
    def save(self, path: str) -> None:
        os.makedirs(path, exist_ok=True)
        meta = {"class": type(self).__name__, "kind": self.kind, "version": "1.0",
                "numFeatures": int(self.weights.numel()), "intercept": self.intercept,
                "threshold": getattr(self, "threshold", None)}
        with open(os.path.join(path, "metadata.json"), "w") as f:
            json.dump(meta, f)
        with open(os.path.join(path, "weights.txt"), "w") as f:
            f.write(" ".join(repr(float(v)) for v in self.weights.detach().cpu().double().tolist()) + "\n")

    @classmethod
    def load(cls, path: str):
        with open(os.path.join(path, "metadata.json")) as f:
            meta = json.load(f)
        with open(os.path.join(path, "weights.txt")) as f:
            w = torch.tensor([float(t) for t in f.read().split()], dtype=torch.float64)
        klass = {c.__name__: c for c in (LinearRegressionModel, LogisticRegressionModel, SVMModel,
                                         RidgeRegressionModel, LassoModel)}.get(meta["class"], cls)
        m = klass(w, meta["intercept"])
        if meta.get("threshold", None) is not None and hasattr(m, "threshold"):
            m.threshold = meta["threshold"]
        elif hasattr(m, "threshold") and "threshold" in meta and meta["threshold"] is None:
            m.threshold = None
        return m

    def __repr__(self):
        return f"{type(self).__name__}(numFeatures={self.weights.numel()}, intercept={self.intercept})"


class LinearRegressionModel(GeneralizedLinearModel):
    kind = "linear"


class RidgeRegressionModel(LinearRegressionModel):
    pass


class LassoModel(LinearRegressionModel):
    pass


class LogisticRegressionModel(GeneralizedLinearModel):
    kind = "logistic"

    def __init__(self, weights, intercept=0.0):
        super().__init__(weights, intercept)
        self.threshold: Optional[float] = 0.5

    def set_threshold(self, t: float):
        self.threshold = float(t)
        return self

    def clear_threshold(self):
        self.threshold = None
        return self

    def predict_point(self, margin):
        m = torch.as_tensor(margin, dtype=torch.float64)
        score = torch.sigmoid(m)
        if self.threshold is None:
            return score if score.ndim else float(score)
        out = (score > self.threshold).to(torch.float64)
        return out if out.ndim else float(out)


class SVMModel(GeneralizedLinearModel):
    kind = "svm"

    def __init__(self, weights, intercept=0.0):
        super().__init__(weights, intercept)
        self.threshold: Optional[float] = 0.0

    def set_threshold(self, t: float):
        self.threshold = float(t)
        return self

    def clear_threshold(self):
        self.threshold = None
        return self

    def predict_point(self, margin):
        m = torch.as_tensor(margin, dtype=torch.float64)
        if self.threshold is None:
            return m if m.ndim else float(m)
        out = (m > self.threshold).to(torch.float64)
        return out if out.ndim else float(out)


# ------------------------------------------------------------------------------------------ algorithm
def _append_bias(shard: Shard) -> Shard:
    d = shard.num_features
    if shard.is_dense:
        from ..data.batch import padded_ld
        X = torch.zeros(shard.num_rows, padded_ld(d + 1), dtype=shard.X.dtype, device=shard.device)
        X[:, :d] = shard.X[:, :d]
        X[:, d] = 1.0
        return Shard(d + 1, shard.y, X, gidx=shard.gidx, partition=shard.partition)
    n = shard.num_rows
    lens = shard.indptr[1:] - shard.indptr[:-1]
    new_indptr = shard.indptr + torch.arange(n + 1, device=shard.device)
    total = int(new_indptr[-1])
    idx = torch.empty(total, dtype=torch.int32, device=shard.device)
    val = torch.empty(total, dtype=shard.values.dtype, device=shard.device)
    row_of = torch.repeat_interleave(torch.arange(n, device=shard.device), lens)
    pos = torch.arange(int(shard.indptr[-1]), device=shard.device) + row_of
    idx[pos] = shard.indices
    val[pos] = shard.values
    last = new_indptr[1:] - 1
    idx[last] = d
    val[last] = 1.0
    return Shard(d + 1, shard.y, None, new_indptr, idx, val, shard.gidx, shard.partition)


def _column_std(data: ShardedDataset, d: int) -> torch.Tensor:
    """Per-feature standard deviation (StandardScaler withStd, no mean shift)."""
    def seq(acc, s: Shard):
        X = s.to_dense().X[:, :d].double()
        part = torch.stack([X.sum(0), (X * X).sum(0)])
        n = torch.tensor(float(s.num_rows), dtype=torch.float64)
        if acc is None:
            return [part.cpu(), n]
        acc[0] += part.cpu()
        acc[1] += n
        return acc

    def comb(a, b):
        if a is None:
            return b
        if b is None:
            return a
        return [a[0] + b[0], a[1] + b[1]]
    acc = data.map_batch(lambda s: s).aggregate(lambda: None, seq, comb)
    n = float(acc[1])
    mean = acc[0][0] / n
    var = (acc[0][1] / n - mean * mean) * (n / max(n - 1.0, 1.0))
    return torch.sqrt(var.clamp_min(0.0))


class GeneralizedLinearAlgorithm:
    model_class = GeneralizedLinearModel

    def __init__(self):
        self.add_intercept = False
        self.validate_data = True
        self.use_feature_scaling = False
        self.num_features = -1
        self.optimizer: GradientDescent = None  # type: ignore[assignment]
        self.validators: List = []

    def set_intercept(self, b: bool):
        self.add_intercept = bool(b)
        return self

    def set_validate_data(self, b: bool):
        self.validate_data = bool(b)
        return self

    def set_feature_scaling(self, b: bool):
        self.use_feature_scaling = bool(b)
        return self

    def create_model(self, weights: torch.Tensor, intercept: float):
        return self.model_class(weights, intercept)

    def run(self, data: ShardedDataset, initial_weights: Optional[torch.Tensor] = None):
        if not data.batch_capable:
            raise TypeError("GLM training needs a batch-capable ShardedDataset (load_libsvm_file / from_arrays / from_shards)")
        first = next(s for s in data.shards() if s is not None)
        d = first.num_features
        self.num_features = d
        dtype, device = first.dtype, first.device
        if getattr(data, "_want_cache", False) is False and getattr(data, "_shards", None) is None:
            import logging
            logging.getLogger("asyncframework_b200.glm").warning(
                "The input data is not directly cached, which may hurt performance if its parent datasets are also uncached.")
        if self.validate_data:
            for v in self.validators:
                if not v(data):
                    raise ValueError("Input validation failed.")
        if initial_weights is None:
            initial_weights = torch.zeros(d, dtype=dtype, device=device)
        scale = None
        train = data
        if self.use_feature_scaling:
            std = _column_std(data, d)
            scale = torch.where(std > 0, 1.0 / std, torch.zeros_like(std)).to(dtype)

            def scale_shard(s: Shard):
                if s.is_dense:
                    X = s.X.clone()
                    X[:, :d] *= scale.to(s.device)
                    return Shard(d, s.y, X, gidx=s.gidx, partition=s.partition)
                return Shard(d, s.y, None, s.indptr, s.indices, s.values * scale.to(s.device)[s.indices.long()],
                             s.gidx, s.partition)
            train = ShardedDataset.from_shards(data.scheduler, [scale_shard(s) for s in data.shards()], "scaled")
        w0 = initial_weights.to(dtype)
        if self.add_intercept:
            train = ShardedDataset.from_shards(train.scheduler, [_append_bias(s) for s in train.shards()], "with_bias")
            w0 = torch.cat([w0, torch.zeros(1, dtype=dtype, device=w0.device)])
        w = self.optimizer.optimize(train, w0)
        intercept = 0.0
        if self.add_intercept:
            intercept = float(w[-1])
            w = w[:-1]
        if scale is not None:
            w = w * scale.to(w.device)
        model = self.create_model(w, intercept)
        self._post_train(model)
        return model

    def _post_train(self, model) -> None:
        pass


def _binary_label_validator(data: ShardedDataset) -> bool:
    for s in data.shards():
        if s is not None and s.num_rows and not bool(((s.y == 0) | (s.y == 1)).all()):
            return False
    return True


class LinearRegressionWithSGD(GeneralizedLinearAlgorithm):
    """``.h`` receives ``optimizer.get_all_weights()`` after ``train`` — the ASYNC hook used by the
    SparkSGDMLLIB driver to plot objective-vs-time (LinearRegression.scala:130,178-184)."""
    model_class = LinearRegressionModel
    h: List[Tuple[int, torch.Tensor]] = []

    def __init__(self, step_size: float = 1.0, num_iterations: int = 100, reg_param: float = 0.0,
                 mini_batch_fraction: float = 1.0, decay: str = "reference"):
        super().__init__()
        self.optimizer = (GradientDescent(LeastSquaresGradient(), SimpleUpdater(decay))
                          .set_step_size(step_size).set_num_iterations(num_iterations)
                          .set_reg_param(reg_param).set_mini_batch_fraction(mini_batch_fraction))

    def _post_train(self, model):
        LinearRegressionWithSGD.h = self.optimizer.get_all_weights()
        model.weight_history = self.optimizer.get_all_weights()
        model.loss_history = list(self.optimizer.loss_history)

    @classmethod
    def train(cls, data, iterations: int = 100, step: float = 1.0, mini_batch_fraction: float = 1.0,
              initial_weights=None, reg_param: float = 0.0, reg_type: Optional[str] = None,
              intercept: bool = False, validate_data: bool = True, convergence_tol: float = 0.001):
        alg = cls(step, iterations, reg_param, mini_batch_fraction)
        if reg_type is not None:
            alg.optimizer.set_updater(make_updater(reg_type))
        alg.optimizer.set_convergence_tol(convergence_tol)
        alg.set_intercept(intercept).set_validate_data(validate_data)
        return alg.run(data, initial_weights)


class RidgeRegressionWithSGD(LinearRegressionWithSGD):
    model_class = RidgeRegressionModel

    def __init__(self, step_size=1.0, num_iterations=100, reg_param=0.01, mini_batch_fraction=1.0):
        super().__init__(step_size, num_iterations, reg_param, mini_batch_fraction)
        self.optimizer.set_updater(SquaredL2Updater())


class LassoWithSGD(LinearRegressionWithSGD):
    model_class = LassoModel

    def __init__(self, step_size=1.0, num_iterations=100, reg_param=0.01, mini_batch_fraction=1.0):
        super().__init__(step_size, num_iterations, reg_param, mini_batch_fraction)
        self.optimizer.set_updater(L1Updater())


class LogisticRegressionWithSGD(GeneralizedLinearAlgorithm):
    model_class = LogisticRegressionModel

    def __init__(self, step_size: float = 1.0, num_iterations: int = 100, reg_param: float = 0.01,
                 mini_batch_fraction: float = 1.0, decay: str = "reference"):
        super().__init__()
        self.optimizer = (GradientDescent(LogisticGradient(), SquaredL2Updater(decay))
                          .set_step_size(step_size).set_num_iterations(num_iterations)
                          .set_reg_param(reg_param).set_mini_batch_fraction(mini_batch_fraction))
        self.validators = [_binary_label_validator]

    def _post_train(self, model):
        model.weight_history = self.optimizer.get_all_weights()
        model.loss_history = list(self.optimizer.loss_history)

    @classmethod
    def train(cls, data, iterations: int = 100, step: float = 1.0, mini_batch_fraction: float = 1.0,
              initial_weights=None, reg_param: float = 0.0, reg_type: Optional[str] = "l2",
              intercept: bool = False, validate_data: bool = True, convergence_tol: float = 0.001):
        alg = cls(step, iterations, reg_param, mini_batch_fraction)
        if reg_type is not None:
            alg.optimizer.set_updater(make_updater(reg_type))
        else:
            alg.optimizer.set_updater(SimpleUpdater())
        alg.optimizer.set_convergence_tol(convergence_tol)
        alg.set_intercept(intercept).set_validate_data(validate_data)
        return alg.run(data, initial_weights)


class LogisticRegressionWithLBFGS(GeneralizedLinearAlgorithm):
    """``LogisticRegressionWithLBFGS`` (mllib/classification/LogisticRegression.scala:281-300, binary case): LogisticGradient +
    SquaredL2Updater under L-BFGS, **feature scaling on** — the one algorithm of the family that turns it on
    (GeneralizedLinearAlgorithm.scala:281) because the quasi-Newton iteration converges much faster on unit-variance
    columns; the weights are mapped back to the original scale afterwards."""
    model_class = LogisticRegressionModel

    def __init__(self, reg_param: float = 0.0, num_corrections: int = 10, convergence_tol: float = 1e-6, max_iterations: int = 100):
        super().__init__()
        from ..optim.lbfgs import LBFGS
        self.optimizer = (LBFGS(LogisticGradient(), SquaredL2Updater()).set_reg_param(reg_param).set_num_corrections(num_corrections)
                          .set_convergence_tol(convergence_tol).set_max_num_iterations(max_iterations))
        self.validators = [_binary_label_validator]
        self.set_feature_scaling(True)

    def _post_train(self, model):
        model.weight_history = self.optimizer.get_all_weights()
        model.loss_history = list(self.optimizer.loss_history)

    @classmethod
    def train(cls, data, iterations: int = 100, initial_weights=None, reg_param: float = 0.0, reg_type: Optional[str] = "l2",
              intercept: bool = False, corrections: int = 10, tolerance: float = 1e-6, validate_data: bool = True):
        """PySpark signature (python/pyspark/mllib/classification.py ``LogisticRegressionWithLBFGS.train``)."""
        alg = cls(reg_param if reg_type is not None else 0.0, corrections, tolerance, iterations)
        alg.set_intercept(intercept).set_validate_data(validate_data)
        return alg.run(data, initial_weights)


class SVMWithSGD(GeneralizedLinearAlgorithm):
    model_class = SVMModel

    def __init__(self, step_size: float = 1.0, num_iterations: int = 100, reg_param: float = 0.01,
                 mini_batch_fraction: float = 1.0):
        super().__init__()
        self.optimizer = (GradientDescent(HingeGradient(), SquaredL2Updater())
                          .set_step_size(step_size).set_num_iterations(num_iterations)
                          .set_reg_param(reg_param).set_mini_batch_fraction(mini_batch_fraction))
        self.validators = [_binary_label_validator]

    @classmethod
    def train(cls, data, iterations: int = 100, step: float = 1.0, reg_param: float = 0.01,
              mini_batch_fraction: float = 1.0, initial_weights=None, reg_type: Optional[str] = "l2",
              intercept: bool = False, validate_data: bool = True, convergence_tol: float = 0.001):
        alg = cls(step, iterations, reg_param, mini_batch_fraction)
        if reg_type is not None:
            alg.optimizer.set_updater(make_updater(reg_type))
        alg.optimizer.set_convergence_tol(convergence_tol)
        alg.set_intercept(intercept).set_validate_data(validate_data)
        return alg.run(data, initial_weights)
