from .glm import (GeneralizedLinearAlgorithm, GeneralizedLinearModel, LassoModel, LassoWithSGD,
                  LinearRegressionModel, LinearRegressionWithSGD, LogisticRegressionModel,
                  LogisticRegressionWithLBFGS, LogisticRegressionWithSGD, RidgeRegressionModel, RidgeRegressionWithSGD, SVMModel, SVMWithSGD)

__all__ = ["GeneralizedLinearAlgorithm", "GeneralizedLinearModel", "LassoModel", "LassoWithSGD",
           "LinearRegressionModel", "LinearRegressionWithSGD", "LogisticRegressionModel",
           "LogisticRegressionWithLBFGS", "LogisticRegressionWithSGD", "RidgeRegressionModel", "RidgeRegressionWithSGD", "SVMModel", "SVMWithSGD"]
