"""Fused action_to_transaction: discrete actions -> transaction units in front of the step.

CPU: the oracle's restatement against the REFERENCE's own `DQN.action_to_transaction` (dqn.py:160-179), through the
committed golden episode tests/golden/actions.npz (made by tests/golden/make_golden_actions.py, which executes the
reference function cut out of the reference file).  GPU: `Env.step_actions` against the oracle, bit-exact ledgers."""
import os

import numpy as np
import pytest

from madigan_b200 import _abi as A
from madigan_b200.environments.data_source import make_params
from oracle.oracle import OracleBatch, OracleEnv

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "actions.npz")


def golden_params(g):
    ds = {f"pair{i}": {"data_source_type": "OUPair", "data_source_config": dict(
        theta=float(g["theta"]), phi=float(g["phi"]), noise=float(g["noise"]))} for i in range(int(g["pairs"]))}
    P, _ = make_params("Composite", ds, required_margin=float(g["required_margin"]),
                       maintenance_margin=float(g["maintenance_margin"]),
                       transaction_cost_rel=float(g["transaction_cost_rel"]), slippage_rel=float(g["slippage_rel"]))
    return P


def test_oracle_action_units_match_reference_function():
    g = np.load(GOLD)
    P = golden_params(g)
    o = OracleEnv(P, construct=False)
    atoms, unit = int(g["action_atoms"]), float(g["unit_size"])
    zi = 0
    o.reset(normals=g["normals"][zi]); zi += 1
    closes = 0
    for t in range(g["actions"].shape[0]):
        a = g["actions"][t]
        closes += int(((a == 0) & (o.ledger != 0.)).sum())
        u = o.action_units(a, atoms, unit)
        assert np.array_equal(u.view(np.int64), g["units"][t].view(np.int64)), f"step {t}: {u} vs {g['units'][t]}"
        out = o.step(u, normals=g["normals"][zi]); zi += 1
        if out["done"]:
            o.reset(normals=g["normals"][zi]); zi += 1
    assert closes > 10  # "action 0 closes an open position" was exercised


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["pairs8", "ou3", "oupair"])  # sources whose prices are bit-exact on the GPU:
# the units are a function of the price, so only there can the ledger be required bit for bit
def test_step_actions_matches_oracle(case):
    import torch
    from test_gpu_parity import CASES, close, compare_step, cpu, make_pair, noise, sync_state_from_oracle
    N, atoms, unit, window = 384, 7, .04, 8
    env, orc, P = make_pair(case, N, window=window, margins=(.2, .25), costs=(.002, 0., .001, 0.))
    exact = not CASES[case][2]
    rng = np.random.default_rng(5)
    nA = P.n_assets
    orc.reset(fill_ticks=1, clear_nstep=False)
    close(cpu(env.t["price"]), orc.state()["price"], rtol=1e-9)
    sync_state_from_oracle(env, orc)
    nz, uz = noise(rng, P, N, ticks=window)
    env.reset(fill_history=True, normals=nz, uniforms=uz)
    orc.reset(fill_ticks=window, normals=nz, uniforms=uz)
    n_close = 0
    for t in range(40):
        acts = rng.integers(0, atoms, size=(N, nA)).astype(np.int8)
        n_close += int(((acts == 0) & (orc.state()["ledger"].T != 0.)).sum())
        nz, uz = noise(rng, P, N)
        env.step_actions(torch.from_numpy(acts), action_atoms=atoms, unit_size=unit, normals=nz, uniforms=uz)
        orc.step_actions(acts, atoms, unit, normals=nz, uniforms=uz)
        compare_step(env, orc, exact, t)
        done = orc.done.astype(bool)
        if done.any():
            nz, uz = noise(rng, P, N, ticks=window)
            env.reset(mask=torch.from_numpy(orc.done.copy()), fill_history=True, normals=nz, uniforms=uz)
            orc.reset(mask=orc.done.copy(), fill_ticks=window, normals=nz, uniforms=uz)
    assert n_close > 100
