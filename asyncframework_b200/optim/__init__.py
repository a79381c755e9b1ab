from .gradient import (AsyncLeastSquaresGradient, Gradient, HingeGradient, LeastSquaresGradient,
                       LogisticGradient, LOSS_IDS, log1p_exp, make_gradient)
from .updater import L1Updater, SimpleUpdater, SquaredL2Updater, Updater, make_updater
from .gradient_descent import GradientDescent, Optimizer, run_mini_batch_sgd
from .lbfgs import LBFGS, run_lbfgs

__all__ = ["AsyncLeastSquaresGradient", "Gradient", "HingeGradient", "LeastSquaresGradient", "LogisticGradient",
           "LOSS_IDS", "log1p_exp", "make_gradient", "L1Updater", "SimpleUpdater", "SquaredL2Updater", "Updater",
           "make_updater", "GradientDescent", "Optimizer", "run_mini_batch_sgd", "LBFGS", "run_lbfgs"]
