"""Golden vectors for the on-device episode tearsheet, produced by the REFERENCE's own
helpers: `_returns`, `_sharpe_of_returns`, `_sortino_of_returns` and `_drawdowns` are cut out of
/root/reference/madigan/utils/metrics.py with `ast` (the module imports numba, which is not in this image; the one
decorated function, `_returns`, is plain Python under its @nb.njit decorator, which is dropped) and evaluated on seeded
episodes the way `test_summary` (metrics.py:83-171) does.  Run in the build container only:
    python tests/golden/make_golden_tearsheet.py
Writes tests/golden/tearsheet.npz: per episode equity, reward, ledger, cost, and the reference's tearsheet fields."""
import ast
import os

import numpy as np
import pandas as pd

REF_FILE = "/root/reference/madigan/utils/metrics.py"
HERE = os.path.dirname(os.path.abspath(__file__))
WANT = ("_returns", "_sharpe_of_returns", "_sortino_of_returns", "_drawdowns")


def reference_functions():
    tree = ast.parse(open(REF_FILE).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANT]
    for f in fns:
        f.decorator_list = []  # @nb.njit(...) on _returns
    ns = {"np": np, "pd": pd}
    exec(compile(ast.Module(body=fns, type_ignores=[]), REF_FILE, "exec"), ns)
    return ns


def main():
    R = reference_functions()
    rng = np.random.default_rng(11)
    nA = 3
    eps = {}
    for i, T in enumerate((400, 75, 31, 12, 650)):
        eq = 1e6 * np.exp(np.cumsum(rng.standard_normal(T) * (.002 if i != 2 else 0.)))  # episode 2: flat equity
        if i == 3:
            eq = 1e6 * (1 + .001 * np.arange(T))  # monotone up: no downside
        reward = rng.standard_normal(T) * .01
        ledger = rng.integers(-2, 3, size=(T, nA)).astype(float) * (rng.random((T, nA)) < .6)
        cost = np.abs(rng.standard_normal((T, nA))) * (rng.random((T, nA)) < .4)
        ts = np.arange(T, dtype=np.int64)
        out = {"nsteps": T, "mean_equity": pd.Series(eq).mean(), "final_equity": eq[-1], "mean_reward": reward.mean(),
               "max_drawdown": R["_drawdowns"](pd.Series(eq)).iloc[0]["valleys"],
               "mean_transaction_cost": cost.mean(), "total_transaction_cost": cost.sum()}
        max_tf = (ts[-1] - ts[0]) // 10  # metrics.py:108
        for j in range(6):
            tf = 2 ** j
            if tf <= max_tf:
                lr = R["_returns"](eq, ts, tf, False, True)
                out[f"equity_returns_offset_{tf}"] = np.nanmean(lr)
                out[f"equity_sharpe_offset_{tf}"] = R["_sharpe_of_returns"](lr)
                out[f"equity_sortino_offset_{tf}"] = R["_sortino_of_returns"](lr)
            else:
                for nm in ("returns", "sharpe", "sortino"):
                    out[f"equity_{nm}_offset_{tf}"] = np.nan
        for a in range(nA):
            out[f"time_spent_in_pos_{a}"] = len(np.where(ledger[:, a] != 0.)[0]) / len(ledger)  # :116-122
        eps[f"ep{i}_equity"] = eq; eps[f"ep{i}_reward"] = reward; eps[f"ep{i}_ledger"] = ledger; eps[f"ep{i}_cost"] = cost
        eps[f"ep{i}_fields"] = np.array(list(out.keys()))
        eps[f"ep{i}_values"] = np.array([float(v) for v in out.values()])
    np.savez(os.path.join(HERE, "tearsheet.npz"), n_episodes=5, **eps)
    print("wrote tearsheet.npz")


if __name__ == "__main__":
    import warnings
    warnings.simplefilter("ignore")
    main()
