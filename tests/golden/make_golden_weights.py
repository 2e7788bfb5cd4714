"""Golden vectors for the fused DDPG action_to_transaction, produced by the REFERENCE's own
function: `DDPG.action_to_transaction` is cut out of /root/reference/madigan/modelling/algorithm/ddpg.py:182-207 with
`ast` (the module itself cannot be imported here: it pulls in the compiled C++ env) and executed on a stand-in `self`
whose `env` exposes the oracle env's ledgerNormedFull / equity / currentPrices.  Run in the build container only:
    python tests/golden/make_golden_weights.py
Writes tests/golden/weights.npz: per step the fp32 target weights (cash first), the units the reference function
returned, and the normals that drove the prices, so that the test can replay the same episode.
The weights are multiples of 1/64 so that their fp32 sum does not depend on torch's reduction order."""
import ast
import os
import sys
import types

import numpy as np
import torch

REF_FILE = "/root/reference/madigan/modelling/algorithm/ddpg.py"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CONFIG = dict(pairs=2, theta=.015, phi=.01, noise=.03, required_margin=.2, maintenance_margin=.25,
              transaction_cost_rel=.002, slippage_rel=.001, steps=160, seed=78)


def reference_function():
    tree = ast.parse(open(REF_FILE).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DDPG")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "action_to_transaction")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np, "torch": torch}
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
    ws, units, normals = [], [], []
    z = rng.standard_normal(nn)
    normals.append(z)
    o.reset(normals=z)
    for t in range(c["steps"]):
        w = (rng.integers(0, 33, size=nA + 1) / 64.).astype(np.float32)
        if t % 37 == 5:
            w[:] = 0.  # the zero-sum branch (ddpg.py:193-194)
        fake = types.SimpleNamespace(
            n_assets=nA + 1,
            env=types.SimpleNamespace(ledgerNormedFull=np.array(o.ledgerNormedFull), equity=o.equity,
                                      currentPrices=o.prices.copy()))
        tr = np.asarray(fn(fake, torch.from_numpy(w).unsqueeze(0)), dtype=np.float64)[0]
        z = rng.standard_normal(nn)
        out = o.step(tr, normals=z)
        ws.append(w); units.append(tr); normals.append(z)
        if out["done"]:
            z = rng.standard_normal(nn)
            normals.append(z)
            o.reset(normals=z)
    np.savez(os.path.join(HERE, "weights.npz"), weights=np.array(ws, dtype=np.float32), units=np.array(units),
             normals=np.array(normals), **{k: np.array(v) for k, v in CONFIG.items()})
    print("wrote weights.npz:", len(ws), "steps,", len(normals) - len(ws) - 1, "resets")


if __name__ == "__main__":
    main()
