from .glm import (GeneralizedLinearAlgorithm, GeneralizedLinearModel, LassoModel, LassoWithSGD,
                  LinearRegressionModel, LinearRegressionWithSGD, LogisticRegressionModel,
                  LogisticRegressionWithLBFGS, LogisticRegressionWithSGD, RidgeRegressionModel, RidgeRegressionWithSGD, SVMModel, SVMWithSGD)

from .evaluation import BinaryClassificationMetrics, RegressionMetrics  # noqa: E402
from .feature import ColumnSummary, StandardScaler, StandardScalerModel, append_bias, col_stats  # noqa: E402

__all__ = ["BinaryClassificationMetrics", "RegressionMetrics", "ColumnSummary", "StandardScaler", "StandardScalerModel", "append_bias", "col_stats", "GeneralizedLinearAlgorithm", "GeneralizedLinearModel", "LassoModel", "LassoWithSGD",
           "LinearRegressionModel", "LinearRegressionWithSGD", "LogisticRegressionModel",
           "LogisticRegressionWithLBFGS", "LogisticRegressionWithSGD", "RidgeRegressionModel", "RidgeRegressionWithSGD", "SVMModel", "SVMWithSGD"]
