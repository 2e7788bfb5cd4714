"""Golden vectors for the agent's log-return reward, produced by the REFERENCE's own
statements: lines 140-164 of /root/reference/madigan/modelling/algorithm/offpolicy_q.py (the body of
`OffPolicyQ.step`'s training loop from `prev_eq = self._env.equity` to the `reduce_rewards` sum) are cut out with
`ast` -- the module itself cannot be imported here, it pulls in the compiled C++ env -- wrapped in a function and
executed on a stand-in `self` whose `_env` is the oracle env (equity, positionValues, step(transaction) returning
the reference's (state, reward, done, info) tuple with info.brokerResponse).  Run in the build container only:
    python tests/golden/make_golden_agent_reward.py
Writes tests/golden/agent_reward.npz: per step the units, the normals that drove the prices, the reset flags and
the rewards the reference fragment returned with reduce_rewards False and True."""
import ast
import os
import sys
import types

import numpy as np

REF_FILE = "/root/reference/madigan/modelling/algorithm/offpolicy_q.py"
FIRST, LAST = 140, 164
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CONFIG = dict(pairs=3, theta=.015, phi=.01, noise=.03, required_margin=.02, maintenance_margin=.25,
              transaction_cost_rel=.02, transaction_cost_abs=.5, slippage_rel=.001, slippage_abs=.002, steps=200,
              seed=123, scale=60_000.)


def reference_fragment():
    """def frag(self, transaction): <the reference's statements, inside a one-trip loop so its `continue` is legal>;
    return reward, done"""
    tree = ast.parse(open(REF_FILE).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "OffPolicyQ")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "step")
    loop = next(n for n in ast.walk(fn) if isinstance(n, ast.While))
    stmts = [s for s in loop.body if FIRST <= s.lineno and s.end_lineno <= LAST]
    assert stmts and isinstance(stmts[0], ast.Assign) and stmts[0].targets[0].id == "prev_eq", "reference moved"
    src = "def frag(self, transaction):\n    for _once in (0,):\n        pass\n    return reward, done\n"
    mod = ast.parse(src)
    mod.body[0].body[0].body = stmts
    ast.fix_missing_locations(mod)
    ns = {"np": np}
    exec(compile(mod, REF_FILE, "exec"), ns)
    return ns["frag"], [ast.get_source_segment(open(REF_FILE).read(), s) for s in stmts]


class EnvShim:
    """What the fragment touches of the reference env, backed by the oracle env."""

    def __init__(self, o):
        self.o = o

    @property
    def equity(self):
        return self.o.equity

    @property
    def positionValues(self):
        return self.o.ledger.copy() * self.o.prices.copy()  # Portfolio.cpp:170-172

    def step(self, transaction):
        out = self.o.step(transaction, normals=self.normals)
        resp = types.SimpleNamespace(transactionUnits=out["transactionUnits"], transactionPrice=out["transactionPrice"],
                                     transactionCost=out["transactionCost"])
        self.last = out
        return None, out["reward"], out["done"], types.SimpleNamespace(dataEnd=False, brokerResponse=resp)


def make_oracle_env():
    from madigan_b200.environments.data_source import make_params
    from oracle.oracle import OracleEnv
    c = CONFIG
    ds = {f"pair{i}": {"data_source_type": "OUPair",
                       "data_source_config": dict(theta=c["theta"], phi=c["phi"], noise=c["noise"])}
          for i in range(c["pairs"])}
    P, _ = make_params("Composite", ds, required_margin=c["required_margin"],
                       maintenance_margin=c["maintenance_margin"], transaction_cost_rel=c["transaction_cost_rel"],
                       transaction_cost_abs=c["transaction_cost_abs"], slippage_rel=c["slippage_rel"],
                       slippage_abs=c["slippage_abs"])
    return OracleEnv(P, construct=False), P


def main():
    frag, source = reference_fragment()
    print("reference statements executed:\n  " + "\n  ".join(s.splitlines()[0] for s in source))
    o, P = make_oracle_env()
    c = CONFIG
    rng = np.random.default_rng(c["seed"])
    nA, nn = P.n_assets, P.n_normals
    shim = EnvShim(o)
    units, normals, resets, r_full, r_red = [], [], [], [], []
    z = rng.standard_normal(nn)
    normals.append(z)
    o.reset(normals=z)
    for t in range(c["steps"]):
        u = rng.integers(-1, 2, size=nA) * c["scale"] / np.abs(o.prices) * rng.uniform(.2, 1.5, size=nA)
        if t % 7 == 3:
            u = u * 500.  # a ruinous order: 2 % cost on ~30 M notional -> the .35 floor of the log
        if t % 9 == 4:
            u = -o.ledger.copy() * rng.choice([1., 2., .5], size=nA)  # closes, reversals, partial closes
        shim.normals = rng.standard_normal(nn)
        # reduce_rewards False and True from the same pre-step state: run the fragment twice on a copy
        import copy
        saved = copy.deepcopy(o.e)
        rew_full, done = frag(types.SimpleNamespace(_env=shim, reduce_rewards=False), u.copy())
        import ctypes as C
        C.memmove(C.byref(o.e), C.byref(saved), C.sizeof(saved))
        rew_red, done2 = frag(types.SimpleNamespace(_env=shim, reduce_rewards=True), u.copy())
        assert done == done2
        units.append(u); normals.append(shim.normals); r_full.append(np.asarray(rew_full, dtype=np.float64))
        r_red.append(np.asarray(rew_red, dtype=np.float64).reshape(1))
        resets.append(bool(done))
        if done:
            z = rng.standard_normal(nn)
            normals.append(z)
            o.reset(normals=z)
    np.savez(os.path.join(HERE, "agent_reward.npz"), units=np.array(units), normals=np.array(normals),
             resets=np.array(resets), reward_full=np.array(r_full), reward_reduced=np.array(r_red),
             **{k: np.array(v) for k, v in CONFIG.items()})
    print("wrote agent_reward.npz:", len(units), "steps,", int(np.sum(resets)), "resets; clamp hits:",
          int((np.array(r_full) == np.log(.35)).sum()))


if __name__ == "__main__":
    main()
