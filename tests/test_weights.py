"""Fused DDPG action_to_transaction: target portfolio weights -> transaction units in front
of the step.  CPU: the oracle's restatement (`orc_weight_units`) against the reference's own function
(modelling/algorithm/ddpg.py:182-207) through the committed golden episode tests/golden/weights.npz (made by
tests/golden/make_golden_weights.py, which executes the reference function cut out of the reference file) -- bit for
bit.  GPU: `Env.step_weights` against the oracle, bit-exact ledgers."""
import os

import numpy as np
import pytest

from madigan_b200.environments.data_source import make_params
from oracle.oracle import OracleEnv

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "weights.npz")


def golden_params(g):
    ds = {f"pair{i}": {"data_source_type": "OUPair", "data_source_config": dict(
        theta=float(g["theta"]), phi=float(g["phi"]), noise=float(g["noise"]))} for i in range(int(g["pairs"]))}
    P, _ = make_params("Composite", ds, required_margin=float(g["required_margin"]),
                       maintenance_margin=float(g["maintenance_margin"]),
                       transaction_cost_rel=float(g["transaction_cost_rel"]), slippage_rel=float(g["slippage_rel"]))
    return P


def test_oracle_weight_units_match_reference_function():
    g = np.load(GOLD)
    P = golden_params(g)
    o = OracleEnv(P, construct=False)
    zi = 0
    o.reset(normals=g["normals"][zi]); zi += 1
    zero_rows = 0
    for t in range(g["weights"].shape[0]):
        w = g["weights"][t]
        zero_rows += int(w.sum() == 0)
        u = o.weight_units(w)
        assert np.array_equal(u.view(np.int64), g["units"][t].view(np.int64)), f"step {t}: {u} vs {g['units'][t]}"
        out = o.step(u, normals=g["normals"][zi]); zi += 1
        if out["done"]:
            o.reset(normals=g["normals"][zi]); zi += 1
    assert zero_rows >= 3  # the zero-sum branch was exercised


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["pairs8", "ou3", "oupair"])  # sources whose prices are bit-exact on the GPU
def test_step_weights_matches_oracle(case):
    import torch
    from test_gpu_parity import CASES, close, compare_step, cpu, make_pair, noise, sync_state_from_oracle
    N, window = 384, 8
    env, orc, P = make_pair(case, N, window=window, margins=(.2, .25), costs=(.002, 0., .001, 0.))
    exact = not CASES[case][2]
    rng = np.random.default_rng(6)
    nA = P.n_assets
    orc.reset(fill_ticks=1, clear_nstep=False)
    close(cpu(env.t["price"]), orc.state()["price"], rtol=1e-9)
    sync_state_from_oracle(env, orc)
    nz, uz = noise(rng, P, N, ticks=window)
    env.reset(fill_history=True, normals=nz, uniforms=uz)
    orc.reset(fill_ticks=window, normals=nz, uniforms=uz)
    traded = 0
    for t in range(40):
        w = rng.random((N, nA + 1)).astype(np.float32)
        w[rng.random(N) < .05] = 0.           # zero-sum rows
        w[:, 0] += rng.random(N).astype(np.float32) * 3  # mostly cash: keeps the portfolios alive
        nz, uz = noise(rng, P, N)
        env.step_weights(torch.from_numpy(w), normals=nz, uniforms=uz)
        orc.step_weights(w, normals=nz, uniforms=uz)
        compare_step(env, orc, exact, t)
        traded += int((cpu(env.t["trans_units"]) != 0).sum())
        done = orc.done.astype(bool)
        if done.any():
            nz, uz = noise(rng, P, N, ticks=window)
            env.reset(mask=torch.from_numpy(orc.done.copy()), fill_history=True, normals=nz, uniforms=uz)
            orc.reset(mask=orc.done.copy(), fill_ticks=window, normals=nz, uniforms=uz)
    assert traded > 1000
