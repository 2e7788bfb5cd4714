"""On-device episode tearsheet (reference utils/metrics.py:83-171 and helpers :334-421).

CPU: the numpy oracle (oracle/py_oracle.py tearsheet) against golden fields computed by the reference's own helpers
(tests/golden/make_golden_tearsheet.py cuts `_returns`, `_sharpe_of_returns`, `_sortino_of_returns`, `_drawdowns` out
of the reference file).  GPU: the streaming kernels (csrc/mdg_tearsheet.cu) through `EpisodeTearsheet` against the
oracle on real episodes of the env, 1e-9 relative (Welford vs two-pass std)."""
import os
import warnings

import numpy as np
import pytest

from oracle import py_oracle

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tearsheet.npz"))


def _close(a, b, tol):
    if np.isnan(a) and np.isnan(b):
        return True
    return abs(a - b) <= tol * max(1.0, abs(a), abs(b))


def test_oracle_matches_reference_helpers():
    for i in range(int(G["n_episodes"])):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = py_oracle.tearsheet(G[f"ep{i}_equity"], G[f"ep{i}_reward"], G[f"ep{i}_ledger"], G[f"ep{i}_cost"])
        for name, want in zip(G[f"ep{i}_fields"], G[f"ep{i}_values"]):
            assert _close(got[str(name)], float(want), 1e-12), (i, name, got[str(name)], want)


@pytest.mark.gpu
def test_gpu_tearsheet_matches_oracle_on_env_episodes():
    import torch
    from madigan_b200.environments import Env
    from madigan_b200.utils.metrics import EpisodeTearsheet
    N, nA, T = 96, 4, 260
    ds = {f"pair{i}": {"data_source_type": "OUPair", "data_source_config": {"theta": .015, "phi": .01, "noise": .03}}
          for i in range(2)}
    env = Env("Composite", 1e6, {"data_source_config": ds}, n_envs=N, window=8, seed=9, device="cuda")
    env.setTransactionCost(.01, 0.)
    env.reset(fill_history=True)
    ts = EpisodeTearsheet(env, n_offsets=6)
    rng = np.random.default_rng(3)
    rec = dict(eq=[], rew=[], led=[], cost=[], done=[])
    for t in range(T):
        u = torch.from_numpy(rng.integers(-1, 2, size=(N, nA)) * 20_000.).double().cuda()
        # the host-side record is taken exactly where the reference takes it: after the step, before any reset
        env.step(u)
        rec["eq"].append(env.equity.cpu().numpy().copy()); rec["rew"].append(env.t["reward"].cpu().numpy().copy())
        rec["led"].append(env.ledger.cpu().numpy().copy()); rec["cost"].append(env.t["trans_cost"].t().cpu().numpy().copy())
        rec["done"].append(env.t["done"].cpu().numpy().copy())
        ts.update()
        env.reset(mask=env.t["done"], fill_history=True)
    S = {k: v.cpu().numpy() for k, v in ts.summary().items()}
    done = np.array(rec["done"])
    n_frozen = 0
    for e in range(N):
        first = int(np.argmax(done[:, e])) + 1 if done[:, e].any() else T   # the step that reported done is the last one
        n_frozen += first < T
        sl = slice(0, first)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = py_oracle.tearsheet(np.array(rec["eq"])[sl, e], np.array(rec["rew"])[sl, e],
                                       np.array(rec["led"])[sl, e], np.array(rec["cost"])[sl, e])
        for name, w in want.items():
            key = name if not name.startswith("time_spent_in_pos_") else \
                f"time_spent_in_pos_{env._asset_names[int(name.rsplit('_', 1)[1])]}"
            assert _close(float(S[key][e]), float(w), 1e-9), (e, name, S[key][e], w)
    assert n_frozen > 5 and n_frozen < N  # both frozen (finished) and still-running episodes were compared
    # re-arm the finished envs: they start a fresh record
    ts.reset(torch.from_numpy(done.any(0)).cuda())
    env.step(torch.zeros((N, nA), dtype=torch.float64, device="cuda"), tearsheet=ts, auto_reset=True)
    S2 = ts.summary()
    assert torch.all(S2["nsteps"][torch.from_numpy(done.any(0)).cuda()] == 1)
