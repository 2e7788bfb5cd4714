"""Streaming reward normalisers (reference environments/reward_normalization.pyx:14-272).

CPU: the oracle (oracle/reward_norm_oracle.c) against golden vectors produced by the reference's own Cython module
(tests/golden/make_golden_reward_norm.py cythonizes the reference .pyx in a temp dir and streams it).
GPU: the CUDA kernel (csrc/mdg_rewardnorm.cu) through the C-ABI / the host classes against the oracle and the goldens.
Tolerance: 1e-9 relative (fp64; the kernel and the oracle keep the reference's operation order)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "reward_norm.npz"))
CASES = [k for k in G.files if k not in ("rewards", "resets")]


def _split(case):
    cls, _, w = case.partition("_w")
    return cls, int(w) if w else 1


def _close(a, b, tol=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    scale = np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    err = np.where(both_nan, 0.0, np.abs(a - b) / scale)
    err = np.where(np.isnan(err), np.inf, err)
    return float(err.max())


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_cython(case):
    from oracle.oracle import OracleRewardNorm
    cls, w = _split(case)
    rewards, resets, want = G["rewards"], G["resets"], G[case]
    T, L = rewards.shape
    orc = OracleRewardNorm(cls, L, w)
    got = np.zeros_like(want)
    for t in range(T):
        got[t] = orc.stream(rewards[t], resets[t])
    assert _close(got, want) < 1e-9, case


def test_factory_and_config_keys():
    from madigan_b200.environments import reward_normalization as rn
    assert rn.SharpeFixedWindow.KIND != rn.NullShaper.KIND
    with pytest.raises(NotImplementedError):
        rn.make_reward_normalizer({"reward_shaper_config": {"reward_shaper": "NoSuchShaper"}})


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_gpu_matches_reference_and_oracle(case):
    import torch
    from madigan_b200.environments import reward_normalization as rn
    from oracle.oracle import OracleRewardNorm
    cls, w = _split(case)
    rewards, resets, want = G["rewards"], G["resets"], G[case]
    T, L = rewards.shape
    sh = getattr(rn, cls)(w, n_envs=L, device="cuda")
    orc = OracleRewardNorm(cls, L, w)
    got = np.zeros_like(want)
    for t in range(T):
        got[t] = sh.stream(torch.from_numpy(rewards[t]).cuda(), torch.from_numpy(resets[t]).cuda()).cpu().numpy()
        o = orc.stream(rewards[t], resets[t])
        assert _close(got[t], o) < 1e-9, (case, t)
    assert _close(got, want) < 1e-9, case


@pytest.mark.gpu
def test_gpu_many_envs_factory_and_masked_reset():
    """65,536 envs x 200 steps against the oracle: every env its own queue; a masked reset touches only its envs."""
    import torch
    from madigan_b200.environments import reward_normalization as rn
    from oracle.oracle import OracleRewardNorm
    N = 65_536
    rng = np.random.default_rng(7)
    for name, conf in (("SharpeFixedWindow", {"reward_shaper": "SharpeFixedWindow", "window": 30}),
                       ("SortinoFixedWindowB", {"reward_shaper": "SortinoFixedWindowB", "window": 12})):
        sh = rn.make_reward_normalizer({"reward_shaper_config": conf}, n_envs=N, device="cuda")
        assert type(sh).__name__ == name
        orc = OracleRewardNorm(name, N, conf["window"])
        for t in range(60):
            r = rng.standard_normal(N) * .01
            m = rng.random(N) < .02 if t else None
            g = sh.stream(torch.from_numpy(r).cuda(), None if m is None else torch.from_numpy(m).cuda()).cpu().numpy()
            assert _close(g, orc.stream(r, m)) < 1e-9, (name, t)
        mask = np.zeros(N, bool); mask[::3] = True
        sh.reset(torch.from_numpy(mask).cuda()); orc.reset(mask)
        r = rng.standard_normal(N) * .01
        assert _close(sh.stream(torch.from_numpy(r).cuda()).cpu().numpy(), orc.stream(r)) < 1e-9
    null = rn.make_reward_normalizer({"reward_shaper_config": {"reward_shaper": None}}, n_envs=8, device="cuda")
    r = torch.arange(8, dtype=torch.float64, device="cuda")
    assert torch.equal(null.stream(r), r)
