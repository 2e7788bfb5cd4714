"""CPU: the oracle's n-step pops (full-buffer pop, drain on done, shaped rewards) against the transitions the
REFERENCE's ReplayBuffer + NStepBuffer held after the seeded golden episode (tests/golden/replay.npz)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from madigan_b200.environments.data_source import make_params, make_reward  # noqa: E402
from oracle.oracle import OracleBatch  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replay.npz")
CASES = {
    "dsr_n5": dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .01}, nstep_return=5),
    "sum_n3": dict(reward_shaper_config={"reward_shaper": "None"}, nstep_return=3),
    "cos_n4": dict(reward_shaper_config={"reward_shaper": "cosine_port_shaper", "desired_portfolio": [1., 0., 0.],
                                         "cosine_temp": .025}, nstep_return=4),
    "ddr_n1": dict(reward_shaper_config={"reward_shaper": "DDR", "adaptation_rate": .001}, nstep_return=1),
}
N, K, T, SEED, SCALE = 6, 8, 90, 2024, 500_000.


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_pops_match_reference_replay_buffer(name):
    rw = CASES[name]
    P, _ = make_params("OUPair", {"theta": .015, "phi": .01, "noise": .03}, required_margin=.1,
                       maintenance_margin=.25, transaction_cost_rel=.02, slippage_rel=.001)
    R = make_reward(rw["reward_shaper_config"], rw["nstep_return"], .99, True, n_assets=2)
    orc = OracleBatch(N, P, R, window=K, seed=7)
    rng = np.random.default_rng(SEED)
    orc.reset(fill_ticks=K, normals=rng.standard_normal((K, P.n_normals, N)), uniforms=rng.random((K, 1, N)))
    rows = [[] for _ in range(N)]
    length = [0] * N
    for t in range(T):
        st = orc.state()
        price = np.where(np.abs(st["price"]) > 1e-9, st["price"], 1.).T
        a = rng.integers(-1, 2, size=(N, 2)).astype(np.float64)
        units = np.ascontiguousarray(a * (SCALE / np.abs(price)) * rng.uniform(.2, 1.5, size=(N, 2)))
        orc.step(units, normals=rng.standard_normal((P.n_normals, N)), uniforms=rng.random((1, N)))
        done = orc.done.astype(bool)
        for e in range(N):
            length[e] += 1
            for k in range(int(orc.n_popped[e])):  # pop k leaves from a buffer of length[e] - k entries
                L = length[e] - k
                rows[e].append([e, t - L, t, orc.shaped_reward[k][0][e], float(done[e])])
            length[e] -= int(orc.n_popped[e])
        if done.any():
            orc.reset(mask=done.copy(), fill_ticks=K, normals=rng.standard_normal((K, P.n_normals, N)),
                      uniforms=rng.random((K, 1, N)))
    got = np.array([r for env_rows in rows for r in env_rows])
    gold = np.load(GOLD)[name]
    assert got.shape[0] == gold.shape[0]
    assert np.array_equal(got[:, :3], gold[:, :3]) and np.array_equal(got[:, 4], gold[:, 4])
    np.testing.assert_allclose(got[:, 3], gold[:, 3], rtol=1e-12, atol=1e-14)
