"""Golden vectors for the fused action_to_transaction, produced by the REFERENCE's own
function: `DQN.action_to_transaction` is cut out of /root/reference/madigan/modelling/algorithm/dqn.py:160-179 with
`ast` (the module itself cannot be imported here: it pulls in the compiled C++ env) and executed on a stand-in `self`
whose `_env` exposes the oracle env's availableMargin / currentPrices / ledger.  Run in the build container only:
    python tests/golden/make_golden_actions.py
Writes tests/golden/actions.npz: per step the actions, the units the reference function returned, and the normals
that drove the prices, so that the test can replay the same episode."""
import ast
import os
import sys
import types
from typing import Union  # noqa: F401  (used by the extracted function's annotations)

import numpy as np
import torch  # noqa: F401

REF_FILE = "/root/reference/madigan/modelling/algorithm/dqn.py"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CONFIG = dict(pairs=2, theta=.015, phi=.01, noise=.03, required_margin=.2, maintenance_margin=.25,
              transaction_cost_rel=.002, slippage_rel=.001, action_atoms=5, unit_size=.05, steps=160, seed=77)


def reference_function():
    tree = ast.parse(open(REF_FILE).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DQN")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "action_to_transaction")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np, "torch": torch, "Union": Union}
    exec(compile(mod, REF_FILE, "exec"), ns)
    return ns["action_to_transaction"]


def make_oracle_env():
    from madigan_b200.environments.data_source import make_params
    from oracle.oracle import OracleEnv
    c = CONFIG
    ds = {f"pair{i}": {"data_source_type": "OUPair",
                       "data_source_config": dict(theta=c["theta"], phi=c["phi"], noise=c["noise"])}
          for i in range(c["pairs"])}
    P, _ = make_params("Composite", ds, required_margin=c["required_margin"],
                       maintenance_margin=c["maintenance_margin"], transaction_cost_rel=c["transaction_cost_rel"],
                       slippage_rel=c["slippage_rel"])
    return OracleEnv(P, construct=False), P


def main():
    fn = reference_function()
    o, P = make_oracle_env()
    c = CONFIG
    rng = np.random.default_rng(c["seed"])
    nA, nn = P.n_assets, P.n_normals
    acts, units, normals = [], [], []
    z = rng.standard_normal(nn)
    normals.append(z)
    o.reset(normals=z)
    for t in range(c["steps"]):
        a = rng.integers(0, c["action_atoms"], size=nA)
        fake = types.SimpleNamespace(
            unit_size=c["unit_size"], action_atoms=c["action_atoms"],
            _env=types.SimpleNamespace(availableMargin=o.availableMargin, currentPrices=o.prices.copy(),
                                       ledger=o.ledger.copy()))
        tr = np.asarray(fn(fake, a), dtype=np.float64)
        z = rng.standard_normal(nn)
        out = o.step(tr, normals=z)
        acts.append(a); units.append(tr); normals.append(z)
        if out["done"]:
            z = rng.standard_normal(nn)
            normals.append(z)
            o.reset(normals=z)
    np.savez(os.path.join(HERE, "actions.npz"), actions=np.array(acts, dtype=np.int8), units=np.array(units),
             normals=np.array(normals), **{k: np.array(v) for k, v in CONFIG.items()})
    print("wrote actions.npz:", len(acts), "steps,", len(normals) - len(acts) - 1, "resets")


if __name__ == "__main__":
    main()
