"""The BASELINE.json configurations (C1-C5) as parity cases: CUDA path through the C-ABI against the
oracle on the same injected noise and actions.  (C5's full-size properties and the bench live in
test_gpu_parity.py::test_roundtrip_property_full_size and bench.py.)"""
import numpy as np
import pytest
import torch

import test_gpu_parity as G
from oracle import py_oracle

pytestmark = pytest.mark.gpu

G.CASES["c1_synth1"] = ("Synth", {"freq": [1.], "mu": [2.], "amp": [1.], "phase": [0.], "dX": .01, "noise": 0.}, True)
G.CASES["c2_ou1"] = ("OU", {"mean": [10.], "theta": [.08], "phi": [.04]}, False)  # madigan/config.yaml:84-98

LOGRET = dict(reward_shaper_config={"reward_shaper": None}, nstep_return=1, reduce_rewards=True)


def test_c1_single_sine_env_log_return_window_64():
    """C1: madigan/config.yaml default shape -- one single-asset sine Synth env, log-return reward, window 64."""
    G.run_case("c1_synth1", N=1, T=200, window=64, reward=LOGRET, margins=(1., .25), scale=30_000.)


def test_c2_4096_single_asset_ou_envs():
    """C2: 4,096 single-asset OU envs (config.yaml's OU), log-return reward, window 64."""
    G.run_case("c2_ou1", N=4096, T=40, window=64, reward=LOGRET, margins=(1., .25), scale=30_000.)


@pytest.mark.parametrize("shaper,n", [("DSR", 1), ("DSR", 20), ("DDR", 1), ("DDR", 20)])
def test_c3_ou_pair_costs_dsr_ddr(shaper, n):
    """C3: OU-pair (2-asset stat-arb) envs with transaction cost .02 + slippage .001, DSR / DDR reward (adaptation
    .001, n-step 1 and 20, discount .99); the oracle bounds the env count of the parity run."""
    G.run_case("oupair", N=2048, T=60, window=64,
               reward=dict(reward_shaper_config={"reward_shaper": shaper, "adaptation_rate": .001}, nstep_return=n,
                           discount=.99, reduce_rewards=True),
               margins=(1., .25), costs=(.02, 0., .001, 0.), scale=60_000.)


def test_c4_composite16_margin_ppc_normalised_window():
    """C4: 16-asset composite portfolios (Synth4 + OU4 + OUPair + SimpleTrend2 + TrendOU2 + TrendyOU2), required
    margin .1 / maintenance .25, cost .001, PPC (cosine) reward n=5, standard_normal fp32 window every step."""
    case, N, k = "composite16", 1024, 64
    reward = dict(reward_shaper_config={"reward_shaper": "cosine_port_shaper", "desired_portfolio": [1.] + [0.] * 16,
                                        "cosine_temp": .025}, nstep_return=5, reduce_rewards=False)
    env, orc, P = G.make_pair(case, N, k, reward, (.1, .25), (.001, 0., 0., 0.))
    rng = np.random.default_rng(44)
    orc.reset(fill_ticks=1, clear_nstep=False)
    G.sync_state_from_oracle(env, orc)
    nz, uz = G.noise(rng, P, N, ticks=k)
    env.reset(fill_history=True, normals=nz, uniforms=uz)
    orc.reset(fill_ticks=k, normals=nz, uniforms=uz)
    seen = set()
    for t in range(30):
        units = G.gen_units(rng, orc, N, 16, 300_000.)
        nz, uz = G.noise(rng, P, N)
        env.step(torch.from_numpy(units), normals=nz, uniforms=uz)
        orc.step(units, normals=nz, uniforms=uz)
        G.compare_step(env, orc, False, t, shaped=True)
        seen |= set(np.unique(orc.risk).tolist())
        got = env.window("standard_normal", dtype=torch.float32).cpu().numpy()
        want = py_oracle.normalise_batch(orc.window(), "standard_normal").astype(np.float32)
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5)
        done = orc.done.astype(bool)
        if done.any():
            nz, uz = G.noise(rng, P, N, ticks=k)
            env.reset(mask=torch.from_numpy(orc.done.copy()), fill_history=True, normals=nz, uniforms=uz)
            orc.reset(mask=orc.done.copy(), fill_ticks=k, normals=nz, uniforms=uz)
    assert len(seen) >= 2  # the margin gates were exercised
