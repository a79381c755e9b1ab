from .glm import (GeneralizedLinearAlgorithm, GeneralizedLinearModel, LassoModel, LassoWithSGD,
                  LinearRegressionModel, LinearRegressionWithSGD, LogisticRegressionModel,
                  LogisticRegressionWithLBFGS, LogisticRegressionWithSGD, RidgeRegressionModel, RidgeRegressionWithSGD, SVMModel, SVMWithSGD)

from .feature import ColumnSummary, StandardScaler, StandardScalerModel, append_bias, col_stats  # noqa: E402

__all__ = ["ColumnSummary", "StandardScaler", "StandardScalerModel", "append_bias", "col_stats", "GeneralizedLinearAlgorithm", "GeneralizedLinearModel", "LassoModel", "LassoWithSGD",
           "LinearRegressionModel", "LinearRegressionWithSGD", "LogisticRegressionModel",
           "LogisticRegressionWithLBFGS", "LogisticRegressionWithSGD", "RidgeRegressionModel", "RidgeRegressionWithSGD", "SVMModel", "SVMWithSGD"]
