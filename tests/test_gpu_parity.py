"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on identical inputs.

Validation mode: both sides get the same pre-generated noise stream and the same
pre-generated transaction units.  The bar (BASELINE.json north_star):
  * ledger positions, transaction units, risk codes, done / marginCall flags: BIT-EXACT;
  * everything that is only +,-,*,/,sqrt in fp64 (prices of OU/OUPair/trend sources, cash,
    meanEntry, borrowedMargin, observation rows): bit-exact as well, because both sides
    are compiled without FMA contraction and fold sums left to right;
  * values that pass through libm transcendentals (log rewards, sin prices, pow in
    DSR/DDR): relative 1e-9 (CUDA's and glibc's log/sin/pow differ in the last ulp).
"""
import os

import numpy as np
import pytest
import torch

from madigan_b200 import _abi as A
from madigan_b200.environments.data_source import make_params, make_reward

pytestmark = pytest.mark.gpu

RTOL = 1e-9

COMPOSITE16 = {
    "sines": {"data_source_type": "Synth", "data_source_config": {
        "freq": [1., 0.3, 2., 0.5], "mu": [2., 2.1, 2.2, 2.3], "amp": [1., 1.2, 1.3, 1.],
        "phase": [0., 1., 2., 1.], "dX": 0.01, "noise": 0.01}},
    "ou": {"data_source_type": "OU", "data_source_config": {
        "mean": [10., 5., 1., 2.], "theta": [.08, .15, .15, .1], "phi": [.04, .04, .04, .04]}},
    "pair": {"data_source_type": "OUPair", "data_source_config": {"theta": .015, "phi": .01, "noise": .03}},
    "trend": {"data_source_type": "SimpleTrend", "data_source_config": {
        "trend_prob": [0.05, 0.02], "min_period": [5, 10], "max_period": [20, 30], "noise": [0.01, 0.005],
        "start": [10., 15.], "dYMin": [0.001, 0.01], "dYMax": [0.003, 0.03]}},
    "trendou": {"data_source_type": "TrendOU", "data_source_config": {
        "trend_prob": [0.05, 0.02], "min_period": [5, 10], "max_period": [20, 30], "dYMin": [0.001, 0.01],
        "dYMax": [0.003, 0.03], "start": [10., 15.], "theta": [.1, .05], "phi": [.02, .01],
        "noise_trend": [.01, .012], "ema_alpha": [0.1, 0.2]}},
    "trendyou": {"data_source_type": "TrendyOU", "data_source_config": {
        "trend_prob": [0.05, 0.02], "min_period": [5, 10], "max_period": [20, 30], "dYMin": [0.001, 0.01],
        "dYMax": [0.003, 0.03], "start": [10., 15.], "theta": [.1, .05], "phi": [.02, .01],
        "noise_trend": [.01, .012], "ema_alpha": [0.1, 0.2]}},
}
PAIRS8 = {f"pair{i}": {"data_source_type": "OUPair",
                       "data_source_config": {"theta": .015, "phi": .01, "noise": .03}} for i in range(8)}

SINE_DYN = {"freqRange": [[.1, 1., .01], [0.3, 3.0, .01], [5., 15., .1], [10., 50., .1]],
            "muRange": [[1., 5., .02], [.3, 3., .05], [.2, 5., .02], [.5, 5., .02]],
            "ampRange": [[1., 5., .01], [.3, 3., .02], [.2, 2., .04], [.5, 5., .05]], "dX": 0.01, "noise": .3}
SINE_TREND = dict(SINE_DYN, trendRange=[[5, 40], [10, 30]], trendIncr=[0.01, 0.02], trendProb=[.02, .05])
SINE_ADDER = {"freq": [1., 0.3, 2., 0.5], "mu": [2., 2.1, 2.2, 2.3], "amp": [1., 1.2, 1.3, 1.],
              "phase": [0., 1., 2., 1.], "dX": 0.01, "noise": .05}
SINE_MIX = {
    "dyn": {"data_source_type": "SineDynamic", "data_source_config": SINE_DYN},
    "pair": {"data_source_type": "OUPair", "data_source_config": {"theta": .015, "phi": .01, "noise": .03}},
    "dyntrend": {"data_source_type": "SineDynamicTrend", "data_source_config": SINE_TREND},
    "trendou": COMPOSITE16["trendou"],
    "dyn2": {"data_source_type": "SineDynamic", "data_source_config": dict(
        freqRange=[[.5, 2., .05], [1., 20., .5], [2., 4., .01]], muRange=[[1., 2., .1]] * 3,
        ampRange=[[.1, .5, .05]] * 3, dX=0.02, noise=0.)},
}

CASES = {
    # name: (data_source_type, data_source_config, has_transcendental_prices)
    "sineadder": ("SineAdder", SINE_ADDER, True),
    "sinedynamic": ("SineDynamic", SINE_DYN, False),       # wave tables come from the host: arithmetic only
    "sinedyntrend": ("SineDynamicTrend", SINE_TREND, False),
    "sine_mix": ("Composite", SINE_MIX, False),
    "synth4": ("Synth", None, True),
    "ou1": ("OU", {"mean": [10.], "theta": [.08], "phi": [.04]}, False),
    "ou3": ("OU", {"mean": [10., 5., 1.], "theta": [.08, .15, .15], "phi": [.04, .02, .01]}, False),
    "oupair": ("OUPair", {"theta": .015, "phi": .01, "noise": .03}, False),
    "pairs8": ("Composite", PAIRS8, False),
    "composite16": ("Composite", COMPOSITE16, True),
    "trends6": ("Composite", {k: COMPOSITE16[k] for k in ("trend", "trendou", "trendyou")}, False),
}


def make_pair(case, N, window=8, reward=None, margins=(1., .25), costs=(0., 0., 0., 0.), seed=7, env_offset=0):
    from madigan_b200.environments import Env
    from oracle.oracle import OracleBatch
    ds_type, ds_cfg, _ = CASES[case]
    P, _ = make_params(ds_type, ds_cfg, required_margin=margins[0], maintenance_margin=margins[1],
                       transaction_cost_rel=costs[0], transaction_cost_abs=costs[1],
                       slippage_rel=costs[2], slippage_abs=costs[3])
    R = None
    if reward is not None:
        R = make_reward(reward.get("reward_shaper_config"), reward.get("nstep_return", 1),
                        reward.get("discount", 0.99), reward.get("reduce_rewards", False), n_assets=P.n_assets)
    env = Env(ds_type, 1_000_000., {"data_source_config": ds_cfg}, n_envs=N, window=window, seed=seed,
              env_offset=env_offset, reward=reward)
    env.setRequiredMargin(margins[0])
    env.setMaintenanceMargin(margins[1])
    env.setTransactionCost(costs[0], costs[1])
    env.setSlippage(costs[2], costs[3])
    orc = OracleBatch(N, P, R, window=window, seed=seed, env_offset=env_offset)
    return env, orc, P


def sync_state_from_oracle(env, orc):
    """Start both sides from bit-identical generator state (sources whose reset() keeps the price,
    e.g. OU, would otherwise carry the last-ulp libm difference of the constructor tick)."""
    st = orc.state()
    for name in ("price", "gstate", "timestamp"):
        env.t[name].copy_(torch.from_numpy(st[name]))
    env.invalidate()


def gen_units(rng, orc, N, nA, scale):
    """a*unit with a in {-1,0,1}; sprinkled with exact closes, reversals, oversize orders and zeros."""
    st = orc.state()
    price = np.where(np.abs(st["price"]) > 1e-9, st["price"], 1.0).T  # (N,nA)
    led = st["ledger"].T
    a = rng.integers(-1, 2, size=(N, nA)).astype(np.float64)
    units = a * (scale / np.abs(price)) * rng.uniform(0.2, 1.5, size=(N, nA))
    r = rng.random((N, nA))
    units = np.where(r < 0.08, -led, units)          # exact close
    units = np.where((r >= 0.08) & (r < 0.14), -2. * led, units)  # reversal
    units = np.where((r >= 0.14) & (r < 0.17), 0.5 * -led, units)  # partial close
    units = np.where((r >= 0.17) & (r < 0.19), units * 50., units)  # oversize -> insufficient margin
    return np.ascontiguousarray(units)


def noise(rng, P, N, ticks=None):
    shape_n = (max(P.n_normals, 1), N) if ticks is None else (ticks, max(P.n_normals, 1), N)
    shape_u = (max(P.n_uniforms, 1), N) if ticks is None else (ticks, max(P.n_uniforms, 1), N)
    return rng.standard_normal(shape_n), rng.random(shape_u)


def same_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return np.array_equal(a.view(np.int64), b.view(np.int64)) or np.array_equal(a, b, equal_nan=True)


def close(a, b, rtol=RTOL, atol=1e-12):
    np.testing.assert_allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol, equal_nan=True)


def cpu(x):
    return x.detach().cpu().numpy()


def compare_windows(env, orc, exact):
    """The (N,k,*) windows an agent would get: ring rows written by steps + the env-major rows of the last
    reset's history fill (prices) / the flat portfolio (weights), against the oracle's rings."""
    nv = orc.n_valid
    w = cpu(env.window(None, n_valid=nv))
    if exact:
        assert same_bits(w, orc.window(nv)), "price window"
    else:
        close(w, orc.window(nv))
    close(cpu(env.portfolio_window(nv)), orc.port_window(nv), rtol=1e-13 if exact else RTOL, atol=1e-15 if exact else 1e-10)
    assert np.array_equal(cpu(env.time_window(nv)), orc.time_window(nv)), "time window"


def compare_step(env, orc, exact_prices, t, shaped=False):
    T = env.t
    tag = f"step {t}"
    assert np.array_equal(cpu(T["ledger"]), orc.state()["ledger"]), tag + " ledger"
    st = orc.state()
    assert np.array_equal(cpu(T["trans_units"]), orc.trans_units), tag + " transactionUnits"
    assert np.array_equal(cpu(T["risk"]), orc.risk), tag + " riskInfo"
    assert np.array_equal(cpu(T["done"]), orc.done), tag + " done"
    assert np.array_equal(cpu(T["margin_call"]), orc.margin_call), tag + " marginCall"
    assert np.array_equal(cpu(T["timestamp"]), st["timestamp"]), tag + " timestamp"
    if exact_prices:
        for name in ("price", "mean_entry", "borrowed", "cash"):
            assert same_bits(cpu(T[name]), st[name]), f"{tag} {name} not bit-exact"
        assert same_bits(cpu(T["gstate"])[:max(1, st["gstate"].shape[0])], st["gstate"]), tag + " gstate"
        assert same_bits(cpu(T["trans_price"]), orc.trans_price), tag + " transactionPrice"
        assert same_bits(cpu(T["trans_cost"]), orc.trans_cost), tag + " transactionCost"
        assert same_bits(cpu(T["obs_price"][env.head]), orc.obs_price[orc.head]), tag + " obs price row"
        # portfolio weights are multiplied by 1/equity in the kernel (observations carry the 1e-9 bar)
        close(cpu(T["obs_port"][env.head]), orc.obs_port[orc.head], rtol=1e-13, atol=1e-15)
    else:
        for name in ("price", "mean_entry", "borrowed", "cash"):
            close(cpu(T[name]), st[name])
        close(cpu(T["trans_price"]), orc.trans_price)
        close(cpu(T["trans_cost"]), orc.trans_cost)
        close(cpu(T["obs_price"][env.head]), orc.obs_price[orc.head])
        close(cpu(T["obs_port"][env.head]), orc.obs_port[orc.head], atol=1e-10)
    close(cpu(T["reward"]), orc.reward)
    if shaped:
        close(cpu(T["agent_reward"]), orc.agent_reward)
        assert np.array_equal(cpu(T["n_popped"]), orc.n_popped), tag + " n_popped"
        npop = orc.n_popped
        g, o = cpu(T["shaped_reward"]), orc.shaped_reward
        for k in range(g.shape[0]):
            m = npop > k
            close(g[k][:, m], o[k][:, m], atol=1e-11)
        close(cpu(T["shaper_A"]), st["shaper_A"], atol=1e-14)
        close(cpu(T["shaper_B"]), st["shaper_B"], atol=1e-16)
        assert np.array_equal(cpu(T["nstep_len"]), st["nstep_len"]), tag + " nstep_len"


def run_case(case, N, T, window=8, reward=None, margins=(1., .25), costs=(0., 0., 0., 0.), scale=30_000.,
             seed=3, reset_done=True):
    env, orc, P = make_pair(case, N, window, reward, margins, costs)
    exact = not CASES[case][2]
    rng = np.random.default_rng(seed)
    nA = P.n_assets
    # the constructor tick used Philox on both sides (Box-Muller through CUDA's vs glibc's libm)
    orc.reset(fill_ticks=1, clear_nstep=False)
    close(cpu(env.t["price"]), orc.state()["price"], rtol=1e-9)
    sync_state_from_oracle(env, orc)
    nz, uz = noise(rng, P, N, ticks=window)
    env.reset(fill_history=True, normals=nz, uniforms=uz)
    orc.reset(fill_ticks=window, normals=nz, uniforms=uz)
    compare_windows(env, orc, exact)
    n_done = 0
    for t in range(T):
        units = gen_units(rng, orc, N, nA, scale)
        nz, uz = noise(rng, P, N)
        env.step(torch.from_numpy(units), normals=nz, uniforms=uz)
        orc.step(units, normals=nz, uniforms=uz)
        compare_step(env, orc, exact, t, shaped=reward is not None)
        done = orc.done.astype(bool)
        n_done += int(done.sum())
        if reset_done and done.any():
            nz, uz = noise(rng, P, N, ticks=window)
            env.reset(mask=torch.from_numpy(orc.done.copy()), fill_history=True, normals=nz, uniforms=uz)
            orc.reset(mask=orc.done.copy(), fill_ticks=window, normals=nz, uniforms=uz)
            assert np.array_equal(cpu(env.t["ledger"]), orc.state()["ledger"])
            compare_windows(env, orc, exact)
    # derived accounting at the end
    d_o = orc.derived()
    d_g = env._derived()
    for name, ref in d_o.items():
        if name == "risk":
            assert np.array_equal(cpu(d_g[name]), ref)
        elif exact:
            assert same_bits(cpu(d_g[name]), ref), name
        else:
            close(cpu(d_g[name]), ref, atol=1e-9)
    return n_done


@pytest.mark.parametrize("case", list(CASES))
def test_step_parity_cash_account(case):
    """required margin 1.0 (no leverage), zero costs: the reference's default test setting."""
    run_case(case, N=257, T=40)


@pytest.mark.parametrize("case", ["ou1", "oupair", "pairs8", "composite16"])
def test_step_parity_margin_costs_slippage(case):
    """leverage 10x, maintenance .25, transaction cost + slippage; oversize orders hit the risk gates,
    volatile paths hit margin calls -> done -> masked reset with history fill."""
    n_done = run_case(case, N=300, T=60, margins=(.1, .25), costs=(.02, 1.5, .001, .002), scale=400_000.)
    assert n_done >= 0


@pytest.mark.parametrize("case,N", [("pairs8", 384), ("oupair", 256), ("pairs8", 128)])
def test_step_parity_whole_blocks_tma_path(case, N):
    """Launches of whole 128-env blocks take the TMA-staged operand path of the all-pairs kernel (3-D tensor copy of the
    state slab, 2-D tensor copy of the units, two-stage ring); the sizes above (257, 300) take the plain-load path.
    One pair (a single ring stage), eight pairs, one block and three blocks; leverage, costs, slippage, resets."""
    run_case(case, N=N, T=50, margins=(.1, .25), costs=(.02, 1.5, .001, .002), scale=400_000.)
    run_case(case, N=N, T=20, reward=dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .001},
                                         nstep_return=1, reduce_rewards=True))


def test_risk_paths_are_exercised():
    env, orc, P = make_pair("oupair", 512, margins=(.1, .25), costs=(.001, 0., 0., 0.))
    orc.reset(fill_ticks=1, clear_nstep=False)
    sync_state_from_oracle(env, orc)
    rng = np.random.default_rng(5)
    seen = set()
    dones = 0
    for t in range(80):
        units = gen_units(rng, orc, 512, 2, 600_000.)
        nz, uz = noise(rng, P, 512)
        env.step(torch.from_numpy(units), normals=nz, uniforms=uz)
        orc.step(units, normals=nz, uniforms=uz)
        assert np.array_equal(cpu(env.t["risk"]), orc.risk)
        assert np.array_equal(cpu(env.t["ledger"]), orc.state()["ledger"])
        seen |= set(np.unique(orc.risk).tolist())
        dones += int(orc.done.sum())
        if orc.done.any():
            env.reset(mask=torch.from_numpy(orc.done.copy()), fill_history=False, normals=nz, uniforms=uz)
            orc.reset(mask=orc.done.copy(), fill_ticks=1, normals=nz[None], uniforms=uz[None])
    assert A.RISK_INSUFF_MARGIN in seen and A.RISK_MARGIN_CALL in seen, seen
    assert dones > 0


@pytest.mark.parametrize("mode", ["hold", "single"])
def test_hold_and_single_asset_steps(mode):
    env, orc, P = make_pair("ou3", 130, margins=(.5, .25))
    orc.reset(fill_ticks=1, clear_nstep=False)
    sync_state_from_oracle(env, orc)
    rng = np.random.default_rng(11)
    for t in range(25):
        nz, uz = noise(rng, P, 130)
        if mode == "hold":
            env.step(normals=nz, uniforms=uz)
            orc.step(None, normals=nz, uniforms=uz)
        else:
            idx = t % 3
            u = rng.integers(-1, 2, size=130) * 20_000. * rng.random(130)
            env.step(idx, torch.from_numpy(u), normals=nz, uniforms=uz)
            orc.step(u, asset_idx=idx, normals=nz, uniforms=uz)
        compare_step(env, orc, True, t)


SHAPERS = [
    dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .001}, nstep_return=1, reduce_rewards=True),
    dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .01}, nstep_return=5, reduce_rewards=False),
    dict(reward_shaper_config={"reward_shaper": "DDR", "adaptation_rate": .001}, nstep_return=4, reduce_rewards=True),
    dict(reward_shaper_config={"reward_shaper": None}, nstep_return=3, reduce_rewards=False),
    dict(reward_shaper_config={"reward_shaper": "cosine_port_shaper", "desired_portfolio": [1., 0., 0.],
                               "cosine_temp": .025}, nstep_return=5, reduce_rewards=False),
    dict(reward_shaper_config={"reward_shaper": "sharpe_shaper"}, nstep_return=6, reduce_rewards=False),
    dict(reward_shaper_config={"reward_shaper": "sortino_shaperA", "sortino_exp": 2.}, nstep_return=6, reduce_rewards=True),
    dict(reward_shaper_config={"reward_shaper": "sortino_shaperB", "sortino_exp": 2.}, nstep_return=4, reduce_rewards=False),
]


@pytest.mark.parametrize("reward", SHAPERS, ids=lambda r: f"{r['reward_shaper_config']['reward_shaper']}-n{r['nstep_return']}")
def test_reward_shapers_in_kernel(reward):
    n_done = run_case("oupair", N=200, T=50, reward=reward, margins=(.1, .25), costs=(.02, 0., .001, 0.),
                      scale=400_000.)
    assert n_done >= 0


def test_headline_shape_dsr_16_assets():
    """C5 shape: 8 OU pairs, DSR n=1, reduced rewards, cost + slippage."""
    run_case("pairs8", N=384, T=30, window=64,
             reward=dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .001},
                         nstep_return=1, reduce_rewards=True),
             margins=(1., .25), costs=(.02, 0., .001, 0.), scale=40_000.)


def test_free_running_philox_matches_oracle():
    """Free-running mode: Philox counters are bit-identical to the oracle's, Box-Muller goes through
    CUDA's log/sincos instead of glibc's -> prices agree to 1e-9 over a short horizon."""
    env, orc, P = make_pair("pairs8", 200, seed=1234, env_offset=1000)
    orc.reset(fill_ticks=1, clear_nstep=False)
    rng = np.random.default_rng(2)
    for t in range(20):
        units = gen_units(rng, orc, 200, 16, 20_000.)
        env.step(torch.from_numpy(units))
        orc.step(units)
        close(cpu(env.t["price"]), orc.state()["price"], rtol=1e-9)
        assert np.array_equal(cpu(env.t["ledger"]), orc.state()["ledger"])


@pytest.mark.parametrize("case", ["sinedynamic", "sine_mix", "sineadder"])
def test_free_running_sine_sources_match_oracle(case):
    """Free-running Philox with the SINE* sources: the constructor and reset() draws of (freq, mu, amp) (streams 3
    and 2), the packed random booleans and the trend draws are bit-identical on both sides; normals go through CUDA's
    log/sincos -> 1e-9 on prices.  Auto-reset on, so the list-driven refill kernel runs these generators too."""
    N = 300
    env, orc, P = make_pair(case, N, window=8, margins=(.1, .25), costs=(.02, 0., .001, 0.), seed=77, env_offset=500)
    orc.reset(fill_ticks=1, clear_nstep=False)
    rng = np.random.default_rng(5)
    nA = P.n_assets
    exact_rows = []  # generator-state rows that no normal draw touches: freq, mu, amp, phasor, flags
    for i in range(nA):
        g = P.gen[i]
        if g.type in (10, 11):
            exact_rows += list(range(g.gslot, g.gslot + 4 * int(g.p[0])))
    n_done = 0
    for t in range(40):
        units = gen_units(rng, orc, N, nA, 300_000.)
        env.step(torch.from_numpy(units), auto_reset=True)
        orc.step(units)
        done = orc.done.astype(bool)
        assert np.array_equal(cpu(env.t["done"]), orc.done)
        n_done += int(done.sum())
        if done.any():  # the CUDA side has already refilled these envs (auto_reset)
            orc.reset(mask=orc.done.copy(), fill_ticks=8)
        st = orc.state()
        assert np.array_equal(cpu(env.t["ledger"]), st["ledger"])
        if exact_rows:
            assert same_bits(cpu(env.t["gstate"])[exact_rows], st["gstate"][exact_rows]), f"step {t} gstate"
        close(cpu(env.t["price"]), st["price"], rtol=1e-9)
    assert n_done > 0 or case == "sineadder"


def test_sharding_is_invisible():
    """Two slabs with env_offset reproduce one big batch bit for bit (Philox uses the global env id)."""
    from madigan_b200.environments import Env
    cfg = {"data_source_config": CASES["pairs8"][1]}
    big = Env("Composite", 1e6, cfg, n_envs=256, window=8, seed=99)
    lo = Env("Composite", 1e6, cfg, n_envs=128, window=8, seed=99, env_offset=0)
    hi = Env("Composite", 1e6, cfg, n_envs=128, window=8, seed=99, env_offset=128)
    g = torch.Generator().manual_seed(0)
    for t in range(10):
        u = (torch.randint(-1, 2, (256, 16), generator=g).double() * 3000.)
        big.step(u)
        lo.step(u[:128].contiguous())
        hi.step(u[128:].contiguous())
    for name in ("price", "ledger", "cash", "gstate"):
        whole = big.t[name]
        parts = torch.cat([lo.t[name], hi.t[name]], dim=-1)
        assert torch.equal(whole, parts), name


def test_roundtrip_property_full_size():
    """Size-independent property at BASELINE size (65,536 envs x 16 assets): buy then sell the same
    units at unchanged prices (noise-free hold) returns every ledger to exactly zero and, with zero
    costs, cash to exactly init_cash; accounting identities of envTest.cpp:212-266 hold."""
    from madigan_b200.environments import Env
    N = 65_536
    cfg = {"data_source_config": {"freq": [1.] * 16, "mu": [5.] * 16, "amp": [0.] * 16, "phase": [0.] * 16,
                                  "dX": 0.01, "noise": 0.}}
    env = Env("Synth", 1e6, cfg, n_envs=N, window=4)
    env.setRequiredMargin(1.)
    env.setMaintenanceMargin(.25)
    g = torch.Generator().manual_seed(1)
    u = torch.randint(1, 1000, (N, 16), generator=g).double()
    env.step(u)
    eq = env.equity
    ident = env.cash + env.assetValue - env.borrowedMargin
    assert torch.allclose(eq, ident, rtol=0, atol=1e-9)
    assert torch.allclose(eq, env.balance + env.pnl + env.usedMargin, rtol=0, atol=1e-6)
    assert torch.equal(env.ledger, u.to(env.device))
    env.step(-u)
    assert torch.count_nonzero(env.ledger) == 0
    assert torch.equal(env.cash, torch.full_like(env.cash, 1e6))
    assert torch.count_nonzero(env.t["borrowed"]) == 0


def test_step_autoreset_single_call_equals_two_calls():
    """mdg_step_autoreset (one host call, optionally on a bound stream) == mdg_step + mdg_reset_ws(mask=done)."""
    from madigan_b200.environments import Env
    cfg = {"data_source_config": PAIRS8}
    rw = dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .001}, nstep_return=3, reduce_rewards=True)
    envs = [Env("Composite", 1e6, cfg, n_envs=3000, window=16, seed=99, reward=rw) for _ in range(3)]
    for e in envs:
        e.setRequiredMargin(.1); e.setTransactionCost(.02, 0.); e.setSlippage(.001, 0.)
        e.reset(fill_history=True)
    side = torch.cuda.Stream()
    envs[2].bind_stream(side)
    g = torch.Generator().manual_seed(3)
    n_done = 0
    for t in range(60):
        units = (torch.randint(-1, 2, (3000, 16), generator=g).double() * 40_000.).cuda()
        envs[0].step(units, auto_reset=True)                      # fused call
        envs[1].step(units)                                       # two calls
        envs[1]._reset_launch(envs[1].t["done"], 16, True, None, None)
        side.wait_stream(torch.cuda.current_stream())
        envs[2].step(units, auto_reset=True)                      # fused call on its own stream
        torch.cuda.current_stream().wait_stream(side)
        n_done += int(envs[0].t["done"].sum())
        for name in ("price", "ledger", "cash", "timestamp", "reset_ts", "obs_price", "pre_price", "shaped_reward",
                     "shaper_A", "nstep_len", "done"):
            a = envs[0].t[name]
            assert torch.equal(a, envs[1].t[name]) and torch.equal(a, envs[2].t[name]), (t, name)
    assert n_done > 20


@pytest.mark.gpu
def test_multi_wave_kernel_variant_parity():
    """Launches above 151,552 envs take the multi-wave instantiation of the step kernel (168-register budget, predicate
    arithmetic in the gate): the same oracle comparison as everywhere else, at 155,648 envs x 16 assets with DSR."""
    N = 155_648
    reward = dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .001}, nstep_return=1, discount=.99,
                  reduce_rewards=True)
    env, orc, P = make_pair("pairs8", N, 8, reward, (1., .25), (.02, 0., .001, 0.))
    rng = np.random.default_rng(1)
    orc.reset(fill_ticks=1, clear_nstep=False)
    sync_state_from_oracle(env, orc)
    nz, uz = noise(rng, P, N, ticks=8)
    env.reset(fill_history=True, normals=nz, uniforms=uz)
    orc.reset(fill_ticks=8, normals=nz, uniforms=uz)
    seen = set()
    for t in range(8):
        units = (rng.integers(-1, 2, size=(N, 16)) * 9000.).astype(np.float64)
        nz, uz = noise(rng, P, N)
        env.step(torch.from_numpy(units), normals=nz, uniforms=uz)
        orc.step(units, normals=nz, uniforms=uz)
        compare_step(env, orc, True, t, shaped=True)
        seen |= set(np.unique(orc.risk).tolist())
        if orc.done.any():
            nz, uz = noise(rng, P, N, ticks=8)
            env.reset(mask=torch.from_numpy(orc.done.copy()), fill_history=True, normals=nz, uniforms=uz)
            orc.reset(mask=orc.done.copy(), fill_ticks=8, normals=nz, uniforms=uz)
    assert len(seen) >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("knob", ["MDG_NO_TMA", "MDG_NO_BULK", "MDG_BS64"])
def test_step_kernel_fallback_variants_parity(knob):
    """The all-pairs step kernel picks its operand path per launch: TMA tensor copies (state tensors equally spaced, the
    layout Env allocates), nine row copies (any 16-byte aligned layout; MDG_NO_TMA=1 forces it), plain loads
    (MDG_NO_BULK=1), and an optional 64-thread-block instantiation (MDG_BS64=1).  The knobs are read once per process,
    so each variant runs the pairs8 / headline oracle comparisons in a child process."""
    import subprocess
    import sys
    env = dict(os.environ, **{knob: "1"})
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-x", "-m", "gpu", "-k",
                        "(pairs8 and (cash_account or margin_costs)) or headline_shape or autoreset_single_call"],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
