#!/usr/bin/env python
"""The reference's regression example (examples/regression/regression.py:75-119) on a swap of imports: BoFire-style
data model -> surrogate_map -> surrogate.fit(experiments DataFrame) -> surrogate.predict(test DataFrame) ->
NLPD / MSE (src/bark/utils/metrics.py:20-39), on TreeFunction data with mixed continuous / categorical inputs
(src/bofire_mixed/benchmarks/tree_function.py).  The only changed lines against the reference script are the imports
and the benchmark construction (bofire itself is not installed here).  Run on a GPU box:

    python scripts/example_regression.py [--num-train 200] [--num-test 100] [--runs 2]
"""
import argparse
import os
import sys
from time import perf_counter

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bark_b200 import synthetic  # noqa: E402
from bark_b200.bofire_compat import BARKSurrogate, Outputs, surrogate_map  # noqa: E402   (was: bofire_mixed.data_models...)
from bark_b200.domain import CategoricalInput, ContinuousInput, Domain, Inputs  # noqa: E402  (was: bofire.data_models...)


def nlpd(mu, var, test_y):  # src/bark/utils/metrics.py:20-33
    return -(-0.5 * np.log(2 * np.pi * var) - 0.5 * (test_y - mu) ** 2 / var).sum(axis=0) / test_y.shape[0]


def mse(mu, test_y):  # :36-39
    return np.mean(np.square(mu - test_y))


class TreeFunctionBenchmark:
    """tree_function.py:60-97 shaped: `.domain`, `.f(X_df, return_complete=True)`."""

    def __init__(self, dim=4, cat_dim=2, num_cat=4, m=20):
        self.fn = synthetic.TreeFunction(dim=dim, cat_dim=cat_dim, num_cat=num_cat, m=m, function_seed=1)
        cats = [f"c{k}" for k in range(num_cat)]
        feats = [ContinuousInput(f"x_{i}", (0.0, 1.0)) for i in range(dim)]
        feats += [CategoricalInput(f"x_{dim + i}", cats) for i in range(cat_dim)]
        self.cats = cats
        self.domain = Domain(inputs=Inputs(feats), outputs=Outputs())

    def sample(self, n, seed):
        rng = np.random.default_rng(seed)
        X = self.fn.sample_inputs(n, rng)
        df = pd.DataFrame({f.key: X[:, i] for i, f in enumerate(self.domain.inputs.get())})
        for f in self.domain.inputs.get():
            if isinstance(f, CategoricalInput):
                df[f.key] = [self.cats[int(v)] for v in df[f.key]]  # label-valued column, as a bofire experiment has
        return df

    def f(self, X_df, return_complete=True, noise=0.1, seed=0):
        from bark_b200.bofire_compat import transform_inputs
        y = self.fn(transform_inputs(self.domain.inputs, X_df)) + noise * np.random.default_rng(seed).standard_normal(len(X_df))
        out = X_df.copy() if return_complete else pd.DataFrame(index=X_df.index)
        out["y"] = y
        return out


def main(seed, num_train, num_test, runs, model_params):
    benchmark = TreeFunctionBenchmark()
    domain = benchmark.domain
    all_metrics = []
    for run_seed in np.random.default_rng(seed).choice(2**32, size=runs, replace=False):
        surrogate = surrogate_map(BARKSurrogate(inputs=domain.inputs, outputs=domain.outputs, **model_params), seed=int(run_seed))
        experiments = benchmark.f(benchmark.sample(num_train, run_seed), return_complete=True, seed=int(run_seed))
        start = perf_counter()
        surrogate.fit(experiments)
        time_taken = perf_counter() - start
        test_experiments = benchmark.f(benchmark.sample(num_test, run_seed + 1), return_complete=True, seed=int(run_seed) + 1)
        pred = surrogate.predict(test_experiments)
        y = test_experiments["y"].to_numpy()
        all_metrics.append([nlpd(pred["y_pred"].to_numpy(), pred["y_sd"].to_numpy() ** 2, y), mse(pred["y_pred"].to_numpy(), y),
                            time_taken])
    return pd.DataFrame(all_metrics, columns=["NLPD", "MSE", "Time"])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-s", "--seed", type=int, default=0)
    ap.add_argument("--num-train", type=int, default=200)
    ap.add_argument("--num-test", type=int, default=100)
    ap.add_argument("--runs", type=int, default=2)
    a = ap.parse_args()
    df = main(a.seed, a.num_train, a.num_test, a.runs, dict(num_chains=4, num_trees=30, warmup_steps=60, num_samples=5))
    print(df.to_string())
    print("baseline MSE of predicting the mean:", "see y variance; a fitted model should be well below it")
