"""Row a16: the agent's log-return reward (offpolicy_q.py:140-164).

CPU: the oracle's restatement (oracle/mdg_oracle.c agent reward) against the REFERENCE's own statements, through
the committed golden episode tests/golden/agent_reward.npz (made by tests/golden/make_golden_agent_reward.py, which
cuts lines 140-164 out of the reference file with `ast` and executes them on the oracle env's equity /
positionValues / broker response).  Covers per-asset and reduced rewards, the .35 floor of the log, closes,
reversals and resets.  GPU: the kernel (which accumulates the reduced reward as ONE log of the product of the
clamped ratios instead of a sum of logs) against the same golden numbers, 1e-9."""
import os

import numpy as np
import pytest

from madigan_b200.environments.data_source import make_params, make_reward
from oracle.oracle import OracleEnv

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "agent_reward.npz")


def golden_params(g):
    ds = {f"pair{i}": {"data_source_type": "OUPair", "data_source_config": dict(
        theta=float(g["theta"]), phi=float(g["phi"]), noise=float(g["noise"]))} for i in range(int(g["pairs"]))}
    kw = dict(required_margin=float(g["required_margin"]), maintenance_margin=float(g["maintenance_margin"]),
              transaction_cost_rel=float(g["transaction_cost_rel"]), transaction_cost_abs=float(g["transaction_cost_abs"]),
              slippage_rel=float(g["slippage_rel"]), slippage_abs=float(g["slippage_abs"]))
    P, _ = make_params("Composite", ds, **kw)
    return P, ds, kw


@pytest.mark.parametrize("reduce", [False, True])
def test_oracle_agent_reward_matches_reference_fragment(reduce):
    g = np.load(GOLD)
    P, _, _ = golden_params(g)
    R = make_reward({"reward_shaper": None}, 1, .99, reduce, n_assets=P.n_assets)  # sum_default, n=1: pops the raw reward
    o = OracleEnv(P, reward=R, construct=False)
    gold = g["reward_reduced"] if reduce else g["reward_full"]
    zi = 0
    o.reset(normals=g["normals"][zi]); zi += 1
    floor_hits = 0
    for t in range(g["units"].shape[0]):
        out = o.step(g["units"][t], normals=g["normals"][zi]); zi += 1
        ra = 1 if reduce else P.n_assets
        np.testing.assert_allclose(out["agent_reward"][:ra], gold[t], rtol=1e-13, atol=1e-15, err_msg=f"step {t}")
        floor_hits += int(np.sum(gold[t] == np.log(.35)))
        assert out["done"] == bool(g["resets"][t])
        if out["done"]:
            o.reset(normals=g["normals"][zi]); zi += 1
    if not reduce:
        assert floor_hits > 5  # np.maximum(reward, .35) was exercised


@pytest.mark.gpu
@pytest.mark.parametrize("reduce", [False, True])
def test_gpu_agent_reward_matches_reference_fragment(reduce):
    import torch
    from madigan_b200.environments import Env
    g = np.load(GOLD)
    P, ds, kw = golden_params(g)
    rw = dict(reward_shaper_config={"reward_shaper": None}, nstep_return=1, reduce_rewards=reduce)
    env = Env("Composite", 1_000_000., {"data_source_config": ds}, n_envs=3, window=4, seed=1, reward=rw)
    env.setRequiredMargin(kw["required_margin"]); env.setMaintenanceMargin(kw["maintenance_margin"])
    env.setTransactionCost(kw["transaction_cost_rel"], kw["transaction_cost_abs"])
    env.setSlippage(kw["slippage_rel"], kw["slippage_abs"])
    gold = g["reward_reduced"] if reduce else g["reward_full"]
    nn = P.n_normals

    def lanes(z):  # the same noise in every lane
        return np.repeat(np.asarray(z, dtype=np.float64).reshape(nn, 1), 3, axis=1)

    zi = 0
    env.reset(normals=lanes(g["normals"][zi])[None], uniforms=np.zeros((1, 1, 3))); zi += 1
    for t in range(g["units"].shape[0]):
        u = np.repeat(g["units"][t][None], 3, axis=0)
        env.step(torch.from_numpy(np.ascontiguousarray(u)), normals=lanes(g["normals"][zi]), uniforms=np.zeros((1, 3)))
        zi += 1
        got = env.agent_reward.cpu().numpy()
        for lane in range(3):
            np.testing.assert_allclose(got[lane], gold[t], rtol=1e-9, atol=1e-12, err_msg=f"step {t}")
        done = bool(env.t["done"][0])
        assert done == bool(g["resets"][t])
        if done:
            env.reset(normals=lanes(g["normals"][zi])[None], uniforms=np.zeros((1, 1, 3))); zi += 1
