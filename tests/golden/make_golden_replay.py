"""Golden transitions for the device replay ingest, produced by the REFERENCE's own
`ReplayBuffer` + `NStepBuffer` (madigan/utils/buffers/replay_buffer.py:68-92, nstep_buffer.py:315-361): one reference
buffer per env, fed step by step with SARSDs whose rewards / dones / portfolio rows come from the CPU oracle env and
whose `state` / `next_state` carry observation indices, with the agent loop's reset handling (offpolicy_q.py:93-99:
env.reset, nstep buffer cleared, history re-initialised).  Run in the build container only:
    python tests/golden/make_golden_replay.py
Writes tests/golden/replay.npz; tests/test_replay.py replays the same seeded episode on the GPU."""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, "/root/reference")
sys.modules.setdefault("rollers", types.ModuleType("rollers"))
sys.modules["rollers"].Roller = object

CASES = {
    "dsr_n5": dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .01}, nstep_return=5),
    "sum_n3": dict(reward_shaper_config={"reward_shaper": "None"}, nstep_return=3),
    "cos_n4": dict(reward_shaper_config={"reward_shaper": "cosine_port_shaper", "desired_portfolio": [1., 0., 0.],
                                         "cosine_temp": .025}, nstep_return=4),
    "ddr_n1": dict(reward_shaper_config={"reward_shaper": "DDR", "adaptation_rate": .001}, nstep_return=1),
}
N, K, T, SEED, SCALE = 6, 8, 90, 2024, 500_000.
DS = ("OUPair", {"theta": .015, "phi": .01, "noise": .03})
MARGINS, COSTS = (.1, .25), (.02, 0., .001, 0.)


def episode_inputs(P, rng, orc):
    """the units of one step (the same generator as tests/test_gpu_parity.gen_units, simplified)"""
    st = orc.state()
    price = np.where(np.abs(st["price"]) > 1e-9, st["price"], 1.).T
    a = rng.integers(-1, 2, size=(N, P.n_assets)).astype(np.float64)
    return np.ascontiguousarray(a * (SCALE / np.abs(price)) * rng.uniform(.2, 1.5, size=(N, P.n_assets)))


def main():
    from madigan.utils.buffers.replay_buffer import ReplayBuffer
    from madigan.utils.data import SARSD, State
    from madigan_b200.environments.data_source import make_params, make_reward
    from oracle.oracle import OracleBatch
    out = {}
    for name, rw in CASES.items():
        P, _ = make_params(DS[0], DS[1], required_margin=MARGINS[0], maintenance_margin=MARGINS[1],
                           transaction_cost_rel=COSTS[0], transaction_cost_abs=COSTS[1], slippage_rel=COSTS[2],
                           slippage_abs=COSTS[3])
        R = make_reward(rw["reward_shaper_config"], rw["nstep_return"], .99, True, n_assets=P.n_assets)
        orc = OracleBatch(N, P, R, window=K, seed=7)
        rng = np.random.default_rng(SEED)
        nz = rng.standard_normal((K, P.n_normals, N))
        orc.reset(fill_ticks=K, normals=nz, uniforms=rng.random((K, 1, N)))
        bufs = [ReplayBuffer(10_000, rw["nstep_return"], .99, rw["reward_shaper_config"]) for _ in range(N)]
        prev_idx = [-1] * N   # observation index each env's next transition starts from
        for t in range(T):
            units = episode_inputs(P, rng, orc)
            orc.step(units, normals=rng.standard_normal((P.n_normals, N)), uniforms=rng.random((1, N)))
            done = orc.done.astype(bool)
            port = orc.obs_port[orc.head]  # (nA+1, N): ledgerNormedFull after the step
            for e in range(N):
                state = State(np.array([prev_idx[e]]), None, None)
                nxt = State(np.array([t]), port[None, :, e].copy(), None)  # cosine reads next_state.portfolio[-1]
                # np.float64, as the agent loop produces it (offpolicy_q.py:162-164 `sum(reward)` of an ndarray); a bare Python
                # float would make `... + EPS` (a float32 scalar) round the DSR denominator to float32 under NumPy >= 2
                bufs[e].add(SARSD(state, units[e].copy(), np.float64(orc.agent_reward[0, e]), nxt, bool(done[e])))
                prev_idx[e] = t
                if done[e]:
                    bufs[e]._nstep_buffer.clear()  # offpolicy_q.py:94 (already drained by add())
            if done.any():
                orc.reset(mask=done.copy(), fill_ticks=K, normals=rng.standard_normal((K, P.n_normals, N)),
                          uniforms=rng.random((K, 1, N)))
        rows = []
        for e in range(N):
            for j in range(bufs[e].filled):
                s = bufs[e]._buffer[j]
                rows.append([e, int(s.state.price[0]), int(s.next_state.price[0]), float(np.asarray(s.reward).ravel()[0]),
                             float(bool(s.done))] + list(np.asarray(s.action, dtype=np.float64)))
        out[name] = np.array(rows)
        print(name, "transitions", len(rows), "dones", int(sum(r[4] for r in rows)))
    np.savez_compressed(os.path.join(HERE, "replay.npz"), **out)


if __name__ == "__main__":
    main()
