from .glm import (GeneralizedLinearAlgorithm, GeneralizedLinearModel, LassoModel, LassoWithSGD,
                  LinearRegressionModel, LinearRegressionWithSGD, LogisticRegressionModel,
                  LogisticRegressionWithSGD, RidgeRegressionModel, RidgeRegressionWithSGD, SVMModel, SVMWithSGD)

__all__ = ["GeneralizedLinearAlgorithm", "GeneralizedLinearModel", "LassoModel", "LassoWithSGD",
           "LinearRegressionModel", "LinearRegressionWithSGD", "LogisticRegressionModel",
           "LogisticRegressionWithSGD", "RidgeRegressionModel", "RidgeRegressionWithSGD", "SVMModel", "SVMWithSGD"]
