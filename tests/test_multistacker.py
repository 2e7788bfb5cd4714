"""MultiStackerDiscrete (reference utils/preprocessor.py:202-288): dilated windows read from
the env-owned ring with a stride, against the reference's own class (tests/golden/multistacker.npz, made by
tests/golden/make_golden_multistacker.py)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multistacker.npz"))
NORMS = [None, "lookback", "lookback_log", "log", "standard_normal", "log_standard_normal"]


def test_host_logic_lengths_follow_reference_deques():
    from madigan_b200.utils.preprocessor import MultiStackerDiscrete, make_preprocessor
    st = MultiStackerDiscrete(6, [1, 2, 4], 3, norm=False)
    want = {1: [1, 2, 3, 4, 5, 6, 6, 6, 6], 2: [1, 1, 2, 2, 3, 3, 4, 4, 5], 4: [1, 1, 1, 1, 2, 2, 2, 2, 3]}
    for c in range(1, 10):
        st._count = c
        for d in (1, 2, 4):
            assert st._len_of(d) == want[d][c - 1]
    assert st.feature_output_shape == (6, 9)
    cfg = {"preprocessor_type": "MultiStackerDiscrete",
           "preprocessor_config": {"window_length": 6, "dilations": [1, 3], "norm": True, "norm_type": "lookback"}}
    assert isinstance(make_preprocessor(cfg, 3), MultiStackerDiscrete)
    with pytest.raises(NotImplementedError):
        MultiStackerDiscrete(6, [1, 2], 3, norm=True, norm_type="expanding")


@pytest.mark.gpu
@pytest.mark.parametrize("norm", NORMS)
def test_multistacker_matches_reference(norm):
    import torch
    from madigan_b200.environments import Env
    from madigan_b200.utils.data import State
    from madigan_b200.utils.preprocessor import MultiStackerDiscrete
    k, dil = int(G["k"]), [int(d) for d in G["dilations"]]
    prices, ports = G["prices"], G["ports"]
    T, nF = prices.shape
    N = 5
    ring = k * max(dil) + 3
    cfg = {"data_source_config": {"mean": [10.] * nF, "theta": [.1] * nF, "phi": [.02] * nF}}
    env = Env("OU", 1e6, cfg, n_envs=N, window=ring, seed=3, device="cuda")
    env.t["reset_ts"].fill_(-(10 ** 9))  # every row counts as written by a step
    st = MultiStackerDiscrete(k, dil, nF, norm=norm is not None, norm_type=norm or "lookback")
    for t in range(T):
        # the row the step kernel would have written: ring slot t % ring, timestamp t (env 0 carries the golden
        # series, the others a scaled copy)
        env.head = t % ring
        scale = torch.arange(1, N + 1, dtype=torch.float64, device="cuda")
        env.t["obs_price"][env.head].copy_(torch.from_numpy(prices[t]).cuda()[:, None] * scale[None, :])
        env.t["obs_port"][env.head].copy_(torch.from_numpy(ports[t]).cuda()[:, None].expand(nF + 1, N))
        env.t["timestamp"].fill_(t)
        env.n_valid = min(ring, t + 1)
        st.stream_state(State(None, None, None, _ring=(env, 0)))
        if t + 1 in G["counts"]:
            cur = st.current_data()
            want = G[f"price_{norm}_{t + 1}"]
            got = cur.price.cpu().numpy()
            np.testing.assert_allclose(got[0], want, rtol=1e-9, atol=1e-12)
            if norm in (None, "log"):  # not scale-free: env e carries (e+1) x the series
                pass
            else:
                np.testing.assert_allclose(got[3], want, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(cur.portfolio.cpu().numpy()[2], G[f"port_{norm}_{t + 1}"], rtol=0, atol=0)
            np.testing.assert_array_equal(cur.timestamp.cpu().numpy()[1], G[f"time_{norm}_{t + 1}"])
