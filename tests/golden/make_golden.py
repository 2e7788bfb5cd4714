"""Generate golden vectors for stages 3-4 of the hot path by running the REFERENCE's own Python code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
It imports, read-only,
    madigan/utils/buffers/nstep_buffer.py   (reward shapers + NStepBuffer)
    madigan/utils/preprocessor.py           (StackerDiscrete + the six normalisers)
feeds them seeded inputs and stores inputs and outputs in tests/golden/*.npz.  The committed
.npz files are what the tests (CPU and GPU) compare against.
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    sys.path.insert(0, REF)
    # `rollers` (3rd party, unpinned, not installed) is only used by RollerDiscrete (preprocessor.py:23,365)
    dummy = types.ModuleType("rollers")
    dummy.Roller = object
    sys.modules.setdefault("rollers", dummy)
    from madigan.utils.buffers import nstep_buffer
    from madigan.utils import preprocessor
    from madigan.utils.data import SARSD, State
    return nstep_buffer, preprocessor, SARSD, State


SHAPER_CASES = [
    ("dsr_n1", {"reward_shaper": "DSR", "adaptation_rate": .001}, 1),
    ("dsr_n5", {"reward_shaper": "DSR", "adaptation_rate": .01}, 5),
    ("ddr_n1", {"reward_shaper": "DDR", "adaptation_rate": .001}, 1),
    ("ddr_n4", {"reward_shaper": "DDR", "adaptation_rate": .005}, 4),
    ("sum_n3", {"reward_shaper": "None"}, 3),
    ("cosine_n5", {"reward_shaper": "cosine_port_shaper", "desired_portfolio": [1., 0., 0., 0.],
                   "cosine_temp": .025}, 5),
    ("sharpe_n1", {"reward_shaper": "sharpe_shaper"}, 1),
    ("sharpe_n6", {"reward_shaper": "sharpe_shaper"}, 6),
    ("sortinoA_n1", {"reward_shaper": "sortino_shaperA", "sortino_exp": 2.}, 1),
    ("sortinoA_n6", {"reward_shaper": "sortino_shaperA", "sortino_exp": 2.}, 6),
    ("sortinoB_n1", {"reward_shaper": "sortino_shaperB", "sortino_exp": 2.}, 1),
    ("sortinoB_n4", {"reward_shaper": "sortino_shaperB", "sortino_exp": 3.}, 4),
]


def run_shaper(nstep_buffer, SARSD, State, cfg, nstep, rewards, ports, dones, discount=0.99):
    """ReplayBuffer.add (reference: utils/buffers/replay_buffer.py:68-80) around the reference's NStepBuffer:
    add; pop once when full; drain on done.  The agent then calls buffer.clear_nstep() on reset
    (offpolicy_q.py:94), which is a no-op after the drain."""
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        nb = nstep_buffer.NStepBuffer(nstep, discount, cfg)
    T = len(rewards)
    out = np.full((T, nstep) + rewards[0].shape, np.nan)
    npop = np.zeros(T, np.int32)
    for t in range(T):
        ns = State(np.zeros((2, 3)), ports[t][None, :].repeat(2, 0), np.zeros(2))
        sarsd = SARSD(None, None, rewards[t].copy(), ns, bool(dones[t]))
        nb.add(sarsd)
        with contextlib.redirect_stdout(io.StringIO()):
            if nb.full():
                out[t, npop[t]] = np.asarray(nb.pop_nstep_sarsd().reward)
                npop[t] += 1
            if sarsd.done:
                while len(nb) > 0:
                    out[t, npop[t]] = np.asarray(nb.pop_nstep_sarsd().reward)
                    npop[t] += 1
    return out, npop


def make_shapers(nstep_buffer, SARSD, State):
    rng = np.random.default_rng(20261018)
    T = 70
    data = {}
    for name, cfg, n in SHAPER_CASES:
        for ra in (1, 3):
            rewards = [np.log(np.maximum(1 + 0.01 * rng.standard_normal(ra), .35)) for _ in range(T)]
            # exact zeros and a run of identical rewards exercise the denom==0 / diff==0 branches
            rewards[7] = np.zeros(ra)
            rewards[20] = rewards[21] = rewards[22] = np.full(ra, 0.003)
            ports = rng.dirichlet(np.ones(4), size=T) * np.where(rng.random((T, 4)) < .2, -1, 1)
            dones = rng.random(T) < .08
            dones[-1] = True
            out, npop = run_shaper(nstep_buffer, SARSD, State, dict(cfg), n, rewards, ports, dones)
            key = f"{name}_ra{ra}"
            data[key + "_rewards"] = np.array(rewards)
            data[key + "_ports"] = ports
            data[key + "_dones"] = dones
            data[key + "_out"] = out
            data[key + "_npop"] = npop
    np.savez_compressed(os.path.join(HERE, "shapers.npz"), **data)
    return len(data)


NORMS = [None, "lookback", "lookback_log", "log", "standard_normal", "log_standard_normal", "expanding"]


def make_preprocessor(preprocessor, State):
    rng = np.random.default_rng(7)
    k, nF, T = 16, 3, 40
    prices = np.abs(10 + np.cumsum(rng.standard_normal((T, nF)) * .3, axis=0)) + .5
    prices[:, 2] = 5.0  # a constant series: std == 0 -> nan_to_num branch
    ports = rng.dirichlet(np.ones(nF + 1), size=T)
    data = {"prices": prices, "ports": ports}
    for norm in NORMS:
        pre = preprocessor.StackerDiscrete(k, nF, norm=norm is not None, norm_type=norm or "lookback")
        outs, lens, steps = [], [], []
        for t in range(T):
            pre.stream_state(State(prices[t], ports[t], np.int64(t + 1)))
            if t in (2, 9, 15, 16, 25, 39):
                cd = pre.current_data()
                w = np.full((k, nF), np.nan)
                w[:len(pre)] = cd.price
                outs.append(w)
                lens.append(len(pre))
                steps.append(t)
                if norm is None:
                    pw = np.full((k, nF + 1), np.nan)
                    pw[:len(pre)] = cd.portfolio
                    data.setdefault("port_windows", []).append(pw)
                    tw = np.full(k, -1, np.int64)
                    tw[:len(pre)] = cd.timestamp
                    data.setdefault("time_windows", []).append(tw)
        data[f"out_{norm}"] = np.array(outs)
        data["lens"] = np.array(lens)
        data["steps"] = np.array(steps)
    data["port_windows"] = np.array(data["port_windows"])
    data["time_windows"] = np.array(data["time_windows"])
    np.savez_compressed(os.path.join(HERE, "preprocessor.npz"), **data)


def make_stackers(preprocessor, State):
    """StackerDiscretePairs (preprocessor.py:295-321) and StackerDiscreteReturns (:324-333)."""
    rng = np.random.default_rng(11)
    k, T = 12, 30
    data = {}
    for name, nF in (("pairs", 2), ("returns", 4)):
        prices = np.abs(10 + np.cumsum(rng.standard_normal((T, nF)) * .3, axis=0)) + .5
        ports = rng.dirichlet(np.ones(nF + 1), size=T)
        data[f"{name}_prices"], data[f"{name}_ports"] = prices, ports
        for norm in NORMS:
            cls = preprocessor.StackerDiscretePairs if name == "pairs" else preprocessor.StackerDiscreteReturns
            pre = cls(k, nF, norm=norm is not None, norm_type=norm or "lookback")
            outs, ports_out, times_out, steps = [], [], [], []
            for t in range(T):
                pre.stream_state(State(prices[t], ports[t], np.int64(t + 1)))
                if t in (11, 12, 20, 29):  # full windows only
                    cd = pre.current_data()
                    outs.append(cd.price); ports_out.append(cd.portfolio); times_out.append(cd.timestamp)
                    steps.append(t)
            data[f"{name}_out_{norm}"] = np.array(outs)
            data[f"{name}_port"] = np.array(ports_out)
            data[f"{name}_time"] = np.array(times_out)
            data["steps"] = np.array(steps)
    np.savez_compressed(os.path.join(HERE, "stackers.npz"), **data)


if __name__ == "__main__":
    nstep_buffer, preprocessor, SARSD, State = import_reference()
    n = make_shapers(nstep_buffer, SARSD, State)
    make_preprocessor(preprocessor, State)
    make_stackers(preprocessor, State)
    print("golden vectors written:", n, "shaper arrays + preprocessor windows")
