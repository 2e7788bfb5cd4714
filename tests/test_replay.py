"""Device-side n-step replay ingest against the REFERENCE's own ReplayBuffer + NStepBuffer:
tests/golden/replay.npz holds, per env, the transitions the reference buffers contained after a seeded 90-step
episode (made by tests/golden/make_golden_replay.py); the same episode is replayed here on the GPU."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replay.npz")

CASES = {
    "dsr_n5": dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .01}, nstep_return=5),
    "sum_n3": dict(reward_shaper_config={"reward_shaper": "None"}, nstep_return=3),
    "cos_n4": dict(reward_shaper_config={"reward_shaper": "cosine_port_shaper", "desired_portfolio": [1., 0., 0.],
                                         "cosine_temp": .025}, nstep_return=4),
    "ddr_n1": dict(reward_shaper_config={"reward_shaper": "DDR", "adaptation_rate": .001}, nstep_return=1),
}
N, K, T, SEED, SCALE = 6, 8, 90, 2024, 500_000.


def run_episode(name, depth=128):
    from madigan_b200.environments import Env
    from madigan_b200.utils.replay import DeviceReplay
    rw = dict(CASES[name], discount=.99, reduce_rewards=True)
    env = Env("OUPair", 1e6, {"data_source_config": {"theta": .015, "phi": .01, "noise": .03}}, n_envs=N, window=K,
              seed=7, reward=rw)
    env.setRequiredMargin(.1); env.setMaintenanceMargin(.25)
    env.setTransactionCost(.02, 0.); env.setSlippage(.001, 0.)
    nn = env.P.n_normals
    rng = np.random.default_rng(SEED)
    env.reset(fill_history=True, normals=rng.standard_normal((K, nn, N)), uniforms=rng.random((K, 1, N)))
    rp = DeviceReplay(env, depth=depth, norm_type="lookback", dtype=torch.float64)
    rp.observe_start()
    windows = {-1: env.window("lookback").cpu().numpy()}
    for t in range(T):
        price = env.t["price"].cpu().numpy().T
        price = np.where(np.abs(price) > 1e-9, price, 1.)
        a = rng.integers(-1, 2, size=(N, 2)).astype(np.float64)
        units = np.ascontiguousarray(a * (SCALE / np.abs(price)) * rng.uniform(.2, 1.5, size=(N, 2)))
        env.step(torch.from_numpy(units), normals=rng.standard_normal((nn, N)), uniforms=rng.random((1, N)))
        done = env.t["done"].cpu().numpy().astype(bool)
        if done.any():
            env.reset(mask=torch.from_numpy(done), fill_history=True, normals=rng.standard_normal((K, nn, N)),
                      uniforms=rng.random((K, 1, N)))
        rp.add(torch.from_numpy(units))
        windows[t] = env.window("lookback").cpu().numpy()
    return env, rp, windows


@pytest.mark.parametrize("name", list(CASES))
def test_replay_ingest_matches_reference_buffers(name):
    gold = np.load(GOLD)[name]
    env, rp, windows = run_episode(name)
    n = len(rp)
    assert n == len(gold)
    t = {k_: v[:n].cpu().numpy() for k_, v in rp.t.items() if k_.startswith("t_")}
    order = np.argsort(t["t_env"], kind="stable")  # per env, append order == the reference buffer's order
    got = np.column_stack([t["t_env"][order], t["t_state_step"][order], t["t_next_slot"][order],
                           t["t_reward"][order, 0], t["t_done"][order], t["t_action"][order]])
    assert np.array_equal(got[:, 0], gold[:, 0]) and np.array_equal(got[:, 1], gold[:, 1])    # env, state index
    assert np.array_equal(got[:, 2], gold[:, 2]) and np.array_equal(got[:, 4], gold[:, 4])    # next index, done
    assert np.array_equal(got[:, 5:], gold[:, 5:])                                             # actions
    np.testing.assert_allclose(got[:, 3], gold[:, 3], rtol=1e-9, atol=1e-11)                   # n-step shaped rewards
    # the stored observations are the windows the agent saw
    for step, w in windows.items():
        np.testing.assert_array_equal(rp.obs_price[step % rp.depth].cpu().numpy(), w)


def test_replay_sample_gathers_stored_records():
    env, rp, _ = run_episode("dsr_n5")
    batch, _none = rp.sample(512)
    idx = rp.last_idx.cpu().numpy()
    n = len(rp)
    assert (idx >= 0).all() and (idx < n).all() and len(np.unique(idx)) > 150
    te, ss, sn = (rp.t[k_].cpu().numpy()[idx] for k_ in ("t_env", "t_state_slot", "t_next_slot"))
    op, opt = rp.obs_price.cpu().numpy(), rp.obs_port.cpu().numpy()
    np.testing.assert_array_equal(batch.state.price.cpu().numpy(), op[ss, te])
    np.testing.assert_array_equal(batch.next_state.price.cpu().numpy(), op[sn, te])
    np.testing.assert_array_equal(batch.state.portfolio.cpu().numpy(), opt[ss, te])
    np.testing.assert_array_equal(batch.next_state.portfolio.cpu().numpy(), opt[sn, te])
    np.testing.assert_array_equal(batch.action.cpu().numpy(), rp.t["t_action"].cpu().numpy()[idx])
    np.testing.assert_array_equal(batch.reward.cpu().numpy(), rp.t["t_reward"].cpu().numpy()[idx])
    np.testing.assert_array_equal(batch.done.cpu().numpy(), rp.t["t_done"].cpu().numpy()[idx].astype(bool))


def test_replay_never_samples_overwritten_observations():
    """depth 16 < episode length: transitions whose state slot has been overwritten are skipped by the sampler."""
    env, rp, _ = run_episode("sum_n3", depth=16)
    rp.sample(256)
    idx = rp.last_idx.cpu().numpy()
    assert (idx >= 0).all()
    steps = rp.t["t_state_step"].cpu().numpy()[idx]
    assert (steps > rp.step_count - rp.depth).all()
