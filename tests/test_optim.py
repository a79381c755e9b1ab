"""Gradient / Updater / GradientDescent / GLM tests in the GradientDescentSuite style
(mllib test GradientDescentSuite.scala:67-183, LinearRegressionSuite.scala:46-154)."""
import math

import numpy as np
import pytest
import torch

from asyncframework_b200.data import LinearDataGenerator, LogisticRegressionDataGenerator, ShardedDataset, generate_gd_input
from asyncframework_b200.data.batch import Shard
from asyncframework_b200.linalg import Vectors
from asyncframework_b200.models import (LassoWithSGD, LinearRegressionModel, LinearRegressionWithSGD, LogisticRegressionModel,
                                        LogisticRegressionWithSGD, RidgeRegressionWithSGD, SVMWithSGD)
from asyncframework_b200.models.glm import GeneralizedLinearModel
from asyncframework_b200.optim import (AsyncLeastSquaresGradient, GradientDescent, HingeGradient, L1Updater,
                                       LeastSquaresGradient, LogisticGradient, SimpleUpdater, SquaredL2Updater,
                                       run_mini_batch_sgd)


@pytest.mark.parametrize("grad", [LeastSquaresGradient(), AsyncLeastSquaresGradient(), LogisticGradient(), HingeGradient()])
def test_gradients_match_autograd(grad):
    torch.manual_seed(0)
    X = torch.randn(50, 6, dtype=torch.float64)
    y = (torch.rand(50) < 0.5).double()
    w = torch.randn(6, dtype=torch.float64, requires_grad=True)
    coeff, loss = grad.coefficients(X @ w, y)
    loss.sum().backward()
    manual = X.t() @ coeff.detach()
    scale = 2.0 if isinstance(grad, AsyncLeastSquaresGradient) else 1.0      # its reported loss is diff² but coeff is diff
    assert torch.allclose(w.grad, manual * scale, atol=1e-9)
    # record-level API agrees with the batch API
    shard = Shard(6, y, torch.cat([X, torch.zeros(50, 2, dtype=torch.float64)], 1))
    g, l, n, _ = grad.batch(shard, w.detach())
    cum = Vectors.zeros(6)
    tot = sum(grad.compute(shard.row(i).features, float(y[i]), Vectors.dense(w.detach()), cum) for i in range(50))
    assert torch.allclose(cum.values, g, atol=1e-9) and abs(tot - float(l)) < 1e-9 and n == 50


def test_logistic_loss_is_stable_for_huge_margins():
    g = LogisticGradient()
    c, l = g.coefficients(torch.tensor([800.0, -800.0], dtype=torch.float64), torch.tensor([1.0, 0.0], dtype=torch.float64))
    assert torch.isfinite(l).all() and l.abs().max() < 1e-9 and c.abs().max() < 1e-9


def test_multinomial_logistic_gradient_matches_autograd():
    K, d = 4, 3
    torch.manual_seed(1)
    x = torch.randn(d, dtype=torch.float64)
    W = torch.randn((K - 1) * d, dtype=torch.float64, requires_grad=True)
    for label in range(K):
        g, loss = LogisticGradient(K).compute(Vectors.dense(x), float(label), Vectors.dense(W.detach()))
        margins = torch.cat([torch.zeros(1, dtype=torch.float64), W.reshape(K - 1, d) @ x])
        ref = -torch.log_softmax(margins, 0)[label]
        (gr,) = torch.autograd.grad(ref, W)
        assert abs(loss - float(ref)) < 1e-9 and torch.allclose(g.values, gr, atol=1e-9)


def test_updaters_fork_schedules():
    w, g = torch.tensor([1.0, -2.0], dtype=torch.float64), torch.tensor([0.5, 0.5], dtype=torch.float64)
    nw, reg = SimpleUpdater().compute(w, g, 1.0, 4, 0.0)
    assert torch.allclose(nw, w - 0.5 * g) and reg == 0.0                      # eta/sqrt(t) kept (Updater.scala:77)
    nw, reg = SquaredL2Updater().compute(w, g, 0.1, 100, 0.5)
    assert torch.allclose(nw, w * (1 - 0.05) - 0.1 * g)                         # CONSTANT step in the fork (:151-158)
    assert reg == pytest.approx(0.25 * float((nw ** 2).sum()))
    nw2, _ = SquaredL2Updater("sqrt").compute(w, g, 0.1, 100, 0.5)             # upstream behaviour available
    assert torch.allclose(nw2, w * (1 - 0.005) - 0.01 * g)
    nw, reg = L1Updater().compute(torch.tensor([0.3, -2.0], dtype=torch.float64), torch.zeros(2, dtype=torch.float64), 1.0, 9, 0.5)
    assert nw.tolist() == [0.0, -1.5] and reg == pytest.approx(0.75)


def test_gd_loss_decreases_and_first_iteration_loss_with_l2(scheduler):
    pts = generate_gd_input(2.0, -1.5, 2000, 42)
    X = np.array([[1.0, float(p.features[0])] for p in pts])
    y = np.array([p.label for p in pts])
    ds = ShardedDataset.from_arrays(scheduler, X, y, 2)
    w0 = torch.zeros(2, dtype=torch.float64)
    _, loss, _ = run_mini_batch_sgd(ds, LogisticGradient(), SimpleUpdater(), 1.0, 10, 0.0, 1.0, w0)
    assert loss[0] > loss[-1] and all(b < a * 1.05 for a, b in zip(loss, loss[1:]))
    # regularisation: loss[0] differs by regVal(initial weights), then both decrease
    w1 = torch.tensor([-1.0, 1.0], dtype=torch.float64)
    _, l0, _ = run_mini_batch_sgd(ds, LogisticGradient(), SquaredL2Updater(), 1.0, 1, 0.0, 1.0, w1)
    _, l1, _ = run_mini_batch_sgd(ds, LogisticGradient(), SquaredL2Updater(), 1.0, 1, 1.0, 1.0, w1)
    assert l1[0] - l0[0] == pytest.approx(0.5 * 1.0 * float((w1 ** 2).sum()), rel=1e-9)


def test_gd_convergence_tolerance_stops_early(scheduler):
    X, y, _ = LinearDataGenerator.generate_linear_arrays(500, 3, eps=0.01, seed=3)
    ds = ShardedDataset.from_arrays(scheduler, X, y, 2)
    w0 = torch.zeros(3, dtype=torch.float64)
    _, loss_tight, _ = run_mini_batch_sgd(ds, LeastSquaresGradient(), SimpleUpdater(), 1.0, 500, 0.0, 1.0, w0, 0.0)
    _, loss_loose, _ = run_mini_batch_sgd(ds, LeastSquaresGradient(), SimpleUpdater(), 1.0, 500, 0.0, 1.0, w0, 0.1)
    assert len(loss_tight) == 500 and len(loss_loose) < 500


def test_linear_regression_with_sgd_history_and_predict(scheduler, tmp_path):
    ds = LinearDataGenerator.generate_linear_dataset(scheduler, 2000, 10, 0.05, 2, seed=7)
    model = LinearRegressionWithSGD.train(ds, iterations=250, step=1.0, mini_batch_fraction=1.0, convergence_tol=0.0)
    assert np.abs(model.weights.numpy() - ds.true_weights).max() < 0.05
    assert len(model.weight_history) == 3 and LinearRegressionWithSGD.h is model.weight_history   # i = 1, 100, 200
    assert all(isinstance(t, int) for t, _ in model.weight_history)
    x = torch.ones(10, dtype=torch.float64)
    assert model.predict(x) == pytest.approx(float(model.weights.sum()))
    assert model.predict(Vectors.dense(x)) == pytest.approx(float(model.weights.sum()))
    model.save(str(tmp_path / "m"))
    back = GeneralizedLinearModel.load(str(tmp_path / "m"))
    assert isinstance(back, LinearRegressionModel) and torch.allclose(back.weights, model.weights.double())
    # intercept support (trailing 1.0 feature split off afterwards)
    X, y, w = LinearDataGenerator.generate_linear_arrays(2000, 4, eps=0.01, seed=5, intercept=2.0)
    dsi = ShardedDataset.from_arrays(scheduler, X, y, 2)
    mi = LinearRegressionWithSGD.train(dsi, iterations=400, step=1.0, intercept=True, convergence_tol=0.0)
    assert abs(mi.intercept - 2.0) < 0.1 and np.abs(mi.weights.numpy() - w).max() < 0.1


def test_logistic_svm_ridge_lasso(scheduler, tmp_path):
    ds = LogisticRegressionDataGenerator.generate_logistic_dataset(scheduler, 2000, 5, 1.5, 2, seed=9)
    m = LogisticRegressionWithSGD.train(ds, iterations=100, step=1.0, reg_param=0.0, intercept=True)
    X = torch.cat([s.X[:, :5] for s in ds.shards()])
    y = torch.cat([s.y for s in ds.shards()])
    acc = float((m.predict(X) == y).double().mean())
    assert acc > 0.85
    m.clear_threshold()
    p = m.predict(X)
    assert float(p.min()) >= 0.0 and float(p.max()) <= 1.0
    m.set_threshold(0.5).save(str(tmp_path / "lr"))
    assert isinstance(GeneralizedLinearModel.load(str(tmp_path / "lr")), LogisticRegressionModel)
    svm = SVMWithSGD.train(ds, iterations=100, step=1.0, reg_param=0.01, intercept=True)
    assert float((svm.predict(X) == y).double().mean()) > 0.85
    bad = ShardedDataset.from_arrays(scheduler, np.random.randn(10, 2), np.arange(10.0), 2)
    with pytest.raises(ValueError, match="validation"):
        LogisticRegressionWithSGD.train(bad, iterations=2)
    lin = LinearDataGenerator.generate_linear_dataset(scheduler, 500, 5, 0.05, 2, seed=1)
    r = RidgeRegressionWithSGD(1.0, 100, 0.1).run(lin)
    l = LassoWithSGD(1.0, 100, 0.1).run(lin)
    plain = LinearRegressionWithSGD(1.0, 100).run(lin)
    assert float(r.weights.norm()) < float(plain.weights.norm()) and float(l.weights.abs().sum()) < float(plain.weights.abs().sum())


def test_feature_scaling_path(scheduler):
    X, y, w = LinearDataGenerator.generate_linear_arrays(1500, 3, eps=0.01, seed=11)
    X = X * np.array([1.0, 10.0, 0.1])
    ds = ShardedDataset.from_arrays(scheduler, X, y, 2)
    alg = LinearRegressionWithSGD(0.5, 600)
    alg.optimizer.set_convergence_tol(0.0)
    alg.set_feature_scaling(True)
    m = alg.run(ds)
    assert np.abs(m.weights.numpy() - w / np.array([1.0, 10.0, 0.1])).max() < 0.2


# ---- L-BFGS (LBFGSSuite scenarios: loss decreasing, agrees with GD, regularisation, convergence tolerance) ----------------
def _logistic_data(sched, n=600, d=5, seed=3):
    import numpy as np
    from asyncframework_b200.data import ShardedDataset
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, d)) * np.array([1.0, 5.0, 0.2, 2.0, 1.0])          # badly scaled columns
    w_true = np.array([1.5, -0.4, 3.0, 0.0, -1.0])
    y = (rng.random(n) < 1 / (1 + np.exp(-(X @ w_true)))).astype(np.float64)
    return ShardedDataset.from_arrays(sched, X, y, 3).cache(), X, y


def test_lbfgs_loss_is_monotone_and_reaches_the_optimum_gd_creeps_towards():
    from asyncframework_b200.core.scheduler import Scheduler
    from asyncframework_b200.optim import LBFGS, GradientDescent, LogisticGradient, SquaredL2Updater, run_lbfgs
    with Scheduler(3) as s:
        data, X, y = _logistic_data(s)
        w0 = torch.zeros(5, dtype=torch.float64)
        opt = LBFGS(LogisticGradient(), SquaredL2Updater()).set_reg_param(0.01).set_max_num_iterations(60).set_convergence_tol(1e-10)
        w = opt.optimize(data, w0)
        h = opt.loss_history
        assert all(b <= a + 1e-12 for a, b in zip(h, h[1:])) and h[-1] < 0.8 * h[0]
        # first-order optimality of the regularised objective: (1/n) X^T (sigmoid(Xw) - y) + lambda w = 0
        Xt, yt = torch.from_numpy(X), torch.from_numpy(y)
        grad = Xt.T @ (torch.sigmoid(Xt @ w) - yt) / len(y) + 0.01 * w
        assert float(grad.abs().max()) < 1e-5
        assert opt.cost_evaluations < 4 * len(h) + 5                                  # the line search rarely backtracks
        # plain GD on the same objective needs far more passes to get equally close
        gd = (GradientDescent(LogisticGradient(), SquaredL2Updater()).set_step_size(0.05).set_num_iterations(len(h))
              .set_reg_param(0.01).set_convergence_tol(0.0))
        w_gd = gd.optimize(data, w0)
        f = lambda v: float(torch.nn.functional.binary_cross_entropy_with_logits(Xt @ v, yt) + 0.005 * (v ** 2).sum())  # noqa: E731
        assert f(w) < f(w_gd) - 1e-4
        w2, hist2 = run_lbfgs(data, LogisticGradient(), SquaredL2Updater(), 10, 1e-10, 60, 0.01, w0)
        assert torch.allclose(w2, w) and hist2 == h
        loose = LBFGS(LogisticGradient(), SquaredL2Updater()).set_convergence_tol(0.1)
        loose.optimize(data, w0)
        assert len(loose.loss_history) < len(h)                                        # tolerance stops it early
        with pytest.raises(ValueError):
            LBFGS(LogisticGradient(), SquaredL2Updater()).set_num_corrections(0)


def test_logistic_regression_with_lbfgs_scales_features_and_predicts():
    from asyncframework_b200.core.scheduler import Scheduler
    from asyncframework_b200.models import LogisticRegressionWithLBFGS, LogisticRegressionWithSGD
    with Scheduler(3) as s:
        data, X, y = _logistic_data(s, n=800)
        m = LogisticRegressionWithLBFGS.train(data, iterations=50, intercept=True)
        acc = float((m.predict(torch.from_numpy(X)) == torch.from_numpy(y)).double().mean())
        m_sgd = LogisticRegressionWithSGD.train(data, iterations=50, step=1.0, intercept=True)
        acc_sgd = float((m_sgd.predict(torch.from_numpy(X)) == torch.from_numpy(y)).double().mean())
        import numpy as np
        oracle = float(((X @ np.array([1.5, -0.4, 3.0, 0.0, -1.0]) > 0) == (y > 0.5)).mean())     # the generating weights
        assert acc >= oracle - 0.02 and acc >= acc_sgd - 0.01
        assert len(m.loss_history) >= 3 and m.loss_history[-1] < m.loss_history[0]
        assert abs(float(m.weights[2])) > abs(float(m.weights[1]))          # back on the ORIGINAL scale: col 2 has the big true weight


def test_col_stats_standard_scaler_and_append_bias_dense_and_csr():
    """MultivariateOnlineSummarizerSuite / StandardScalerSuite at the behaviour level, on dense and CSR shards."""
    import numpy as np
    from asyncframework_b200.core.scheduler import Scheduler
    from asyncframework_b200.data import ShardedDataset
    from asyncframework_b200.models import StandardScaler, append_bias, col_stats
    rng = np.random.default_rng(2)
    X = rng.standard_normal((500, 6)) * np.array([1, 10, 0.1, 3, 0, 2]) * (rng.random((500, 6)) < 0.6)
    X[:, 4] = 0.0                                                          # a constant column: std 0 -> scaled to 0, not NaN
    y = rng.standard_normal(500)
    with Scheduler(3) as s:
        dense = ShardedDataset.from_arrays(s, X, y, 3)
        sparse = ShardedDataset.from_shards(s, [sh.to_csr() for sh in dense.shards()])
        for ds in (dense, sparse):
            st = col_stats(ds)
            assert st.count == 500 and np.allclose(st.mean.numpy(), X.mean(0)) and np.allclose(st.variance.numpy(), X.var(0, ddof=1))
            assert np.allclose(st.num_nonzeros.numpy(), (X != 0).sum(0)) and np.allclose(st.max.numpy(), X.max(0))
            assert np.allclose(st.min.numpy(), X.min(0)) and np.allclose(st.norm_l1.numpy(), np.abs(X).sum(0))
            assert np.allclose(st.norm_l2.numpy(), np.sqrt((X * X).sum(0)))
            model = StandardScaler().fit(ds)
            scaled = model.transform(ds)
            Z = torch.cat([sh.to_dense().X[:, :6] for sh in scaled.shards()]).numpy()
            order = np.concatenate([np.arange(p, 500, 3) for p in range(3)])                  # from_arrays is round-robin
            std = X.std(0, ddof=1)
            want = np.where(std > 0, X / np.where(std > 0, std, 1), 0.0)[order]
            assert np.allclose(Z, want) and not np.isnan(Z).any()
            assert all(sh.is_dense == orig.is_dense for sh, orig in zip(scaled.shards(), ds.shards()))   # sparsity preserved
            centred = model.set_with_mean(True).transform(ds)
            Zc = torch.cat([sh.X[:, :6] for sh in centred.shards()]).numpy()
            assert np.allclose(Zc.mean(0), 0.0, atol=1e-10)
            model.set_with_mean(False)
            assert np.allclose(model.transform(torch.from_numpy(X[0])).numpy(), want[list(order).index(0)])
            wb = append_bias(ds)
            sh0 = wb.shards()[0]
            assert sh0.num_features == 7 and bool((sh0.to_dense().X[:, 6] == 1.0).all())


def test_regression_and_binary_classification_metrics_match_sklearn():
    import numpy as np
    from sklearn import metrics as skm
    from asyncframework_b200.core.scheduler import Scheduler
    from asyncframework_b200.data import ShardedDataset
    from asyncframework_b200.models import BinaryClassificationMetrics, RegressionMetrics
    rng = np.random.default_rng(4)
    y = rng.standard_normal(300)
    p = y + 0.3 * rng.standard_normal(300)
    with Scheduler(2) as s:
        rm = RegressionMetrics(ShardedDataset.from_records(s, list(zip(p.tolist(), y.tolist())), 3))
    assert np.isclose(rm.mean_squared_error, skm.mean_squared_error(y, p)) and np.isclose(rm.mean_absolute_error, skm.mean_absolute_error(y, p))
    assert np.isclose(rm.r2, skm.r2_score(y, p)) and np.isclose(rm.root_mean_squared_error, np.sqrt(skm.mean_squared_error(y, p)))
    labels = (rng.random(400) < 0.4).astype(float)
    scores = np.round(0.35 * labels + 0.65 * rng.random(400), 2)                  # rounded: ties between classes exist
    bm = BinaryClassificationMetrics(torch.from_numpy(scores), torch.from_numpy(labels))
    assert np.isclose(bm.area_under_roc, skm.roc_auc_score(labels, scores))
    prec, rec, _ = skm.precision_recall_curve(labels, scores)
    assert abs(bm.area_under_pr - skm.auc(rec, prec)) < 0.01                        # same curve, different end-point convention
    assert bm.roc()[0] == (0.0, 0.0) and bm.roc()[-1] == (1.0, 1.0) and len(bm.thresholds()) == len(set(scores.tolist()))
    best_t, best_f = max(bm.f_measure_by_threshold(), key=lambda tf: tf[1])
    assert np.isclose(best_f, skm.f1_score(labels, scores >= best_t))
