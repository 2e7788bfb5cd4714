"""The reference's own ledger / risk known-answer tests, replayed THROUGH THE CUDA PATH (C-ABI, mdg_step).

tests/test_oracle_ledger_kat.py replays madigan/environments/cpp/tests/envTest.py against the CPU oracle; here
the same sequences run on the GPU.  The batched env only exposes Env.step, so:
  * the Synth source runs with dX = 0 and noise = 0 (amp = 0 where a price is prescribed): its tick reproduces the
    same price every step, i.e. a step is exactly one Broker.handleTransaction at a known price;
  * a price change (envTest.py:527,534) is a change of the generator's `mu` followed by a hold step;
  * probes that the reference makes with Portfolio.checkRisk(i, units) on ONE portfolio run in separate LANES of
    the batch (same state, different units), because a green probe executes on the GPU.
Every step is also compared bit for bit with the oracle stepped on the same inputs.  The risk test's units sit
exactly at the availableMargin threshold and one unit either side (envTest.py:521-523): that is the knife-edge on
which the kernel's running-sum gate must hand over to exact_gate (mdg_step_kernel.cuh).
"""
import ctypes as C

import numpy as np
import pytest
import torch
from numpy.testing import assert_allclose

from madigan_b200 import _abi as A
from madigan_b200.environments.data_source import make_params

pytestmark = pytest.mark.gpu

# envTest.py:11-21 with the phase frozen (dX = 0): prices stay at their first-tick values
SYNTH_FROZEN = {'freq': [1., 0.3, 2., 0.5], 'mu': [2., 2.1, 2.2, 2.3], 'amp': [1., 1.2, 1.3, 1.0],
                'phase': [0., 1.0, 2., 1.], 'dX': 0., 'noise': 0.}
INIT_CASH = 1_000_000.


def cpu(x):
    return x.detach().cpu().numpy()


class Pair:
    """madigan_b200.Env (GPU) and OracleBatch (CPU) stepped side by side, compared bit for bit."""

    def __init__(self, N, cfg, required_margin, maintenance_margin, flags=0):
        from madigan_b200.environments import Env
        from oracle.oracle import OracleBatch
        self.N = N
        self.P, _ = make_params("Synth", cfg, init_cash=INIT_CASH, required_margin=required_margin,
                                maintenance_margin=maintenance_margin)
        self.env = Env("Synth", INIT_CASH, {"data_source_config": cfg}, n_envs=N, window=4, seed=1)
        self.env.setRequiredMargin(required_margin)
        self.env.setMaintenanceMargin(maintenance_margin)
        self.env.flags = flags
        self.orc = OracleBatch(N, self.P, None, window=4, seed=1)
        self.orc.reset(fill_ticks=1, clear_nstep=False)  # the constructor tick (Env.h:160)
        self.nA = self.P.n_assets
        self.check("construct")

    def set_mu(self, asset, mu):
        self.env.P.gen[asset].p[1] = mu
        for i in range(self.N):
            self.orc.L.orc_batch_env(self.orc.h, i).contents.P.gen[asset].p[1] = mu

    def step(self, units=None):
        z = np.zeros((max(self.P.n_normals, 1), self.N))
        u = np.zeros((max(self.P.n_uniforms, 1), self.N))
        if units is None:
            self.env.step(normals=z, uniforms=u)
            self.orc.step(None, normals=z, uniforms=u)
            self.check("hold")  # a hold step has no broker response (Env.h:189-204): the tensors keep the last one
            return
        else:
            units = np.ascontiguousarray(units, dtype=np.float64).reshape(self.N, self.nA)
            self.env.step(torch.from_numpy(units), normals=z, uniforms=u)
            self.orc.step(units, normals=z, uniforms=u)
        self.check("step")

    def check(self, tag):
        T, st = self.env.t, self.orc.state()
        for name in ("ledger", "mean_entry", "borrowed", "cash", "price"):
            assert np.array_equal(cpu(T[name]), st[name]), f"{tag}: {name} differs from the oracle"
        if tag == "step":
            assert np.array_equal(cpu(T["risk"]), self.orc.risk), "riskInfo"
            assert np.array_equal(cpu(T["trans_units"]), self.orc.trans_units), "transactionUnits"
            assert np.array_equal(cpu(T["margin_call"]), self.orc.margin_call), "marginCall"
        if tag != "construct":
            assert np.array_equal(cpu(T["done"]), self.orc.done), "done"


def ref_transaction(units, init_cash, price, margin=1.):
    """envTest.py:102-115, verbatim arithmetic."""
    cash = init_cash
    cost = margin * (price * units)
    cash -= cost
    borrowed_margin = (1 - margin) * (price * units)
    if borrowed_margin < 0.:
        cash -= borrowed_margin
        borrowed_margin = 0.
    equity = cash + units * price - borrowed_margin
    return cash, borrowed_margin, equity


@pytest.mark.parametrize("flags", [0, A.FLAG_FORCE_EXACT_GATE], ids=["fast_gate", "forced_exact_gate"])
@pytest.mark.parametrize("margin", [1., .1])
def test_port_accounting_logic(margin, flags):
    """envTest.py:119-145, :255-281, :303-330: +-1000 units at required margin 1.0 / 0.1 -- exact equality."""
    p = Pair(2, SYNTH_FROZEN, margin, .25, flags)
    price = float(cpu(p.env.currentPrices)[0, 0])
    u = np.zeros((2, 4))
    u[:, 0] = [1000., -1000.]
    p.step(u)
    for lane, units in enumerate((1000., -1000.)):
        cash, bm, eq = ref_transaction(units, INIT_CASH, price, margin)
        assert cash == float(p.env.cash[lane])
        assert bm == float(p.env.borrowedMargin[lane])
        assert eq == float(p.env.equity[lane])
        assert int(p.env.t["risk"][0, lane]) == A.RISK_GREEN


@pytest.mark.parametrize("flags", [0, A.FLAG_FORCE_EXACT_GATE], ids=["fast_gate", "forced_exact_gate"])
def test_successive_accounting(flags):
    """envTest.py:404-440 (lane 0), :443-456 (lane 1), :459-472 (lane 2): round trips and reversals at margin 0.1."""
    p = Pair(3, SYNTH_FROZEN, .1, .25, flags)
    e = p.env
    p0 = float(cpu(e.currentPrices)[0, 0])
    seq = np.array([[10_000, 10_000, -20_000, -20_000, 10_000, 10_000],
                    [-10_000, 20_000, 0, 0, 0, 0],
                    [10_000, -20_000, 0, 0, 0, 0]], dtype=np.float64)

    def do(t):
        u = np.zeros((3, 4))
        u[:, 0] = seq[:, t]
        p.step(u)

    do(0)
    assert float(e.assetValue[0]) == p0 * 10_000
    do(1)
    # lane 0: long 20k
    assert float(e.cash[0]) == INIT_CASH - 0.1 * p0 * 20_000
    assert float(e.assetValue[0]) == p0 * 20_000
    assert float(e.usedMargin[0]) == 0.1 * p0 * 20_000
    assert float(e.borrowedMargin[0]) == 0.9 * p0 * 20_000
    assert float(e.borrowedAssetValue[0]) == 0.
    # lane 1: short 10k then buy 20k -> +10k (envTest.py:443-456)
    assert float(e.cash[1]) == INIT_CASH - 0.1 * p0 * 10_000
    assert float(e.assetValue[1]) == p0 * 10_000
    assert float(e.usedMargin[1]) == 0.1 * p0 * 10_000
    assert float(e.borrowedMargin[1]) == 0.9 * p0 * 10_000
    assert float(e.borrowedAssetValue[1]) == 0.
    # lane 2: long 10k then sell 20k -> -10k (envTest.py:459-472)
    assert float(e.cash[2]) == INIT_CASH + p0 * 10_000
    assert float(e.assetValue[2]) == -p0 * 10_000
    assert float(e.usedMargin[2]) == 0.1 * p0 * 10_000
    assert float(e.borrowedMargin[2]) == 0.
    assert float(e.borrowedAssetValue[2]) == -p0 * 10_000
    do(2)  # lane 0 flat again
    for name in ("assetValue", "usedMargin", "borrowedMargin", "borrowedAssetValue"):
        assert_allclose(float(getattr(e, name)[0]), 0., atol=1e-9)
    assert_allclose(float(e.cash[0]), INIT_CASH, rtol=1e-12)
    do(3)  # lane 0 short 20k
    assert_allclose(float(e.cash[0]), INIT_CASH + p0 * 20_000., rtol=1e-12)
    assert_allclose(float(e.assetValue[0]), p0 * -20_000, rtol=1e-12)
    assert_allclose(float(e.usedMargin[0]), 0.1 * p0 * 20_000, rtol=1e-12)
    assert_allclose(float(e.borrowedMargin[0]), 0.)
    assert_allclose(float(e.borrowedAssetValue[0]), p0 * -20_000, rtol=1e-12)
    do(4)
    do(5)
    assert_allclose(float(e.cash[0]), INIT_CASH, rtol=1e-12)
    for name in ("assetValue", "usedMargin", "borrowedMargin", "borrowedAssetValue"):
        assert_allclose(float(getattr(e, name)[0]), 0., atol=1e-9)


def test_multiasset_accounting():
    """envTest.py:475-509: long asset 0, short asset 3 -- as two steps (lane 0) and as ONE step (lane 1): the
    sequential-over-assets semantics of Broker.cpp:144-158 make them identical."""
    p = Pair(2, SYNTH_FROZEN, .1, .25)
    e = p.env
    prices = cpu(e.currentPrices)[0].copy()
    u = np.zeros((2, 4))
    u[0, 0] = 20_000
    u[1, 0], u[1, 3] = 20_000, -20_000
    p.step(u)
    assert float(e.cash[0]) == INIT_CASH - 0.1 * prices[0] * 20_000
    assert float(e.assetValue[0]) == prices[0] * 20_000
    assert float(e.usedMargin[0]) == 0.1 * prices[0] * 20_000
    assert float(e.borrowedMargin[0]) == 0.9 * prices[0] * 20_000
    assert float(e.borrowedAssetValue[0]) == 0.
    u = np.zeros((2, 4))
    u[0, 3] = -20_000
    p.step(u)
    cash = INIT_CASH - (0.1 * prices[0] * 20_000) + (prices[3] * 20_000)
    balance = INIT_CASH - (0.1 * prices[0] * 20_000)
    for lane in (0, 1):
        assert_allclose(float(e.cash[lane]), cash, rtol=1e-12)
        assert_allclose(float(e.balance[lane]), balance, rtol=1e-12)
        assert_allclose(float(e.assetValue[lane]), prices[0] * 20_000 + prices[3] * -20_000, rtol=1e-12)
        assert_allclose(float(e.usedMargin[lane]), 0.1 * prices[0] * 20_000 + 0.1 * prices[3] * 20_000, rtol=1e-12)
        assert_allclose(float(e.borrowedMargin[lane]), 0.9 * prices[0] * 20_000, rtol=1e-12)
        assert_allclose(float(e.borrowedAssetValue[lane]), prices[3] * -20_000, rtol=1e-12)
    for name in ("ledger", "cash", "mean_entry", "borrowed"):
        t = cpu(e.t[name])
        assert np.array_equal(t[..., 0], t[..., 1]), name


@pytest.mark.parametrize("flags", [0, A.FLAG_FORCE_EXACT_GATE], ids=["fast_gate", "forced_exact_gate"])
def test_port_risk_handling(flags):
    """envTest.py:512-547: risk thresholds at exactly availableMargin and +-1 unit, then margin-call prices
    4 -> 3.71 (green) -> 3.69 (margin call).  Lanes: 0/1/2 the three threshold probes; 3 the 3.71 probe;
    4 the 3.69 probe; 5 closes at 3.69; 6, 7 untouched."""
    cfg = {'freq': [1.] * 4, 'mu': [2., 4., 2.2, 2.3], 'amp': [0.] * 4, 'phase': [0.] * 4, 'dX': 0., 'noise': 0.}
    N, reqM = 8, 0.1
    p = Pair(N, cfg, reqM, 1., flags)
    e = p.env
    price = 4.
    assert np.all(cpu(e.currentPrices)[:, 1] == price)
    u = np.zeros((N, 4))
    u[:, 1] = 1_000_000
    p.step(u)
    assert np.all(cpu(e.t["risk"])[1] == A.RISK_GREEN) and np.all(cpu(e.ledger)[:, 1] == 1_000_000)
    am = (float(e.balance[0]) + float(e.pnl[0])) / reqM
    u = np.zeros((N, 4))
    u[0, 1] = (-1. + am) / price
    u[1, 1] = (0. + am) / price   # exactly the threshold: availableMargin <= |amount| -> insufficient
    u[2, 1] = (1. + am) / price
    p.step(u)
    risk = cpu(e.t["risk"])[1]
    assert risk[0] == A.RISK_GREEN and risk[1] == A.RISK_INSUFF_MARGIN and risk[2] == A.RISK_INSUFF_MARGIN
    led = cpu(e.ledger)[:, 1]
    assert led[0] == 1_000_000 + u[0, 1] and led[1] == 1_000_000 and led[2] == 1_000_000
    # price 4 -> 3.71: still green (envTest.py:527-530)
    p.set_mu(1, 3.71)
    p.step(None)
    assert np.all(cpu(e.currentPrices)[:, 1] == 3.71)
    assert np.all(cpu(e.checkRisk())[3:] == A.RISK_GREEN)
    assert not cpu(e.t["done"])[3:].any()
    u = np.zeros((N, 4))
    u[3, 1] = 1_000_000 / price
    p.step(u)
    assert cpu(e.t["risk"])[1, 3] == A.RISK_GREEN and cpu(e.ledger)[3, 1] == 1_000_000 + 250_000
    # price -> 3.69: margin call (envTest.py:534-538)
    p.set_mu(1, 3.69)
    p.step(None)
    assert np.all(cpu(e.checkRisk())[4:] == A.RISK_MARGIN_CALL)
    assert cpu(e.t["done"])[4:].all()  # Env::step: done on a margin call (Env.h:196-198)
    am = (float(e.balance[4]) + float(e.pnl[4])) / reqM
    loss = 1_000_000 * (price - 3.69)
    assert_allclose(-loss, float(e.pnl[4]), rtol=1e-12)
    u = np.zeros((N, 4))
    u[4, 1] = (-1. + am) / price
    u[5, 1] = -1_000_000
    p.step(u)
    assert cpu(e.t["risk"])[1, 4] == A.RISK_MARGIN_CALL and cpu(e.ledger)[4, 1] == 1_000_000
    assert cpu(e.t["risk"])[1, 5] == A.RISK_GREEN and cpu(e.ledger)[5, 1] == 0.
    equity = INIT_CASH - loss
    assert_allclose(equity, float(e.equity[5]), rtol=1e-12)
    assert_allclose(equity, float(e.cash[5]), rtol=1e-12)


def test_broker_risk_handling():
    """envTest.py:550-566: response fields of a green transaction."""
    cfg = {'freq': [1.] * 4, 'mu': [2., 4., 2.2, 2.3], 'amp': [0.] * 4, 'phase': [0.] * 4, 'dX': 0., 'noise': 0.}
    p = Pair(1, cfg, .1, 1.)
    u = np.zeros((1, 4))
    u[0, 1] = 1_000_000
    p.step(u)
    resp = p.env._info.brokerResponse
    assert float(resp.transactionPrice[0, 1]) == 4.
    assert float(resp.transactionCost[0, 1]) == 0.
    assert int(resp.riskInfo[0, 1]) == A.RISK_GREEN


# ---------------------------------------------------------------------------------------------------------------
# knife-edge on a LATER asset: the gate of asset i sees the trades of assets < i of the same step, where the
# kernel's running sums are rounded differently from the oracle's fresh left-to-right folds
# ---------------------------------------------------------------------------------------------------------------
DELTAS = [0., 1e-16, -1e-16, 2.3e-16, -2.3e-16, 1e-14, -1e-14, 1e-13, -1e-13, 1e-12, -1e-12, 1e-10, -1e-10, 1e-8, -1e-8]


@pytest.mark.parametrize("flags", [0, A.FLAG_FORCE_EXACT_GATE], ids=["fast_gate", "forced_exact_gate"])
def test_knife_edge_on_later_assets(flags):
    from madigan_b200.environments import Env
    from oracle.oracle import OracleBatch, OrcEnv, lib as olib
    pairs = {f"pair{i}": {"data_source_type": "OUPair", "data_source_config": {"theta": .015, "phi": .01, "noise": .03}}
             for i in range(8)}
    N, nA, reqM = 16 * len(DELTAS), 16, .2
    costs = dict(transaction_cost_rel=.002, transaction_cost_abs=0.5, slippage_rel=.001, slippage_abs=.002)
    P, _ = make_params("Composite", pairs, required_margin=reqM, maintenance_margin=.25, **costs)
    env = Env("Composite", INIT_CASH, {"data_source_config": pairs}, n_envs=N, window=4, seed=5)
    env.setRequiredMargin(reqM); env.setMaintenanceMargin(.25)
    env.setTransactionCost(.002, .5); env.setSlippage(.001, .002)
    env.flags = flags
    orc = OracleBatch(N, P, None, window=4, seed=5)
    orc.reset(fill_ticks=1, clear_nstep=False)
    st = orc.state()
    for name in ("price", "gstate", "timestamp"):
        env.t[name].copy_(torch.from_numpy(st[name]))
    env.invalidate()
    rng = np.random.default_rng(17)
    L = olib()
    n_insuff = n_green = 0
    for t in range(12):
        st = orc.state()
        price, led = st["price"].T, st["ledger"].T  # (N, nA)
        units = rng.integers(-1, 2, size=(N, nA)) * (20_000. / price) * rng.uniform(.2, 1.5, size=(N, nA))
        for e_ in range(N):
            i = e_ % nA
            delta = DELTAS[e_ // nA]
            # the oracle's portfolio after the trades of assets < i: the exact threshold asset i's gate will see
            scratch = OrcEnv()
            C.memmove(C.byref(scratch), L.orc_batch_env(orc.h, e_), C.sizeof(OrcEnv))
            tp, tu, tc = C.c_double(), C.c_double(), C.c_double()
            for j in range(i):
                L.orc_broker_transaction(C.byref(scratch), j, float(units[e_, j]), C.byref(tp), C.byref(tu), C.byref(tc))
            am = L.orc_available_margin(C.byref(scratch))
            sign = 1. if led[e_, i] >= 0 else -1.  # same side as the position: not an offsetting order
            units[e_, i] = sign * (am / price[e_, i]) * (1. + delta)
        units = np.ascontiguousarray(units)
        z = rng.standard_normal((P.n_normals, N))
        u = np.zeros((1, N))
        env.step(torch.from_numpy(units), normals=z, uniforms=u)
        orc.step(units, normals=z, uniforms=u)
        assert np.array_equal(cpu(env.t["risk"]), orc.risk), f"step {t}: riskInfo"
        assert np.array_equal(cpu(env.t["trans_units"]), orc.trans_units), f"step {t}: transactionUnits"
        s2 = orc.state()
        for name in ("ledger", "mean_entry", "borrowed", "cash", "price"):
            assert np.array_equal(cpu(env.t[name]), s2[name]), f"step {t}: {name}"
        assert np.array_equal(cpu(env.t["done"]), orc.done)
        probe = orc.risk[np.arange(N) % nA, np.arange(N)]
        n_insuff += int((probe == A.RISK_INSUFF_MARGIN).sum())
        n_green += int((probe == A.RISK_GREEN).sum())
        if orc.done.any():
            zz = rng.standard_normal((1, P.n_normals, N))
            env.reset(mask=torch.from_numpy(orc.done.copy()), normals=zz, uniforms=np.zeros((1, 1, N)))
            orc.reset(mask=orc.done.copy(), fill_ticks=1, normals=zz, uniforms=np.zeros((1, 1, N)))
    assert n_insuff > 100 and n_green > 100, (n_insuff, n_green)  # both sides of the threshold were hit


def test_forced_exact_gate_equals_fast_gate_random_orders():
    """Every gate through exact_gate (MDG_FLAG_FORCE_EXACT_GATE) vs the running-sum gate: identical ledgers, risk
    codes and cash over a random-order episode with leverage, costs and slippage (both are also oracle-checked in
    test_gpu_parity.py; this pins them to EACH OTHER at a size the oracle tests do not reach)."""
    from madigan_b200.environments import Env
    pairs = {f"pair{i}": {"data_source_type": "OUPair", "data_source_config": {"theta": .015, "phi": .01, "noise": .03}}
             for i in range(8)}
    N = 8192
    envs = []
    for flags in (0, A.FLAG_FORCE_EXACT_GATE):
        e = Env("Composite", INIT_CASH, {"data_source_config": pairs}, n_envs=N, window=8, seed=21)
        e.setRequiredMargin(.1); e.setMaintenanceMargin(.25); e.setTransactionCost(.02, 1.5); e.setSlippage(.001, .002)
        e.flags = flags
        e.reset(fill_history=True)
        envs.append(e)
    g = torch.Generator().manual_seed(9)
    seen = set()
    for t in range(40):
        a = torch.randint(-1, 2, (N, 16), generator=g).double() * 60_000. * torch.rand((N, 16), generator=g, dtype=torch.float64)
        a = a.cuda()
        for e in envs:
            e.step(a, auto_reset=True)
        for name in ("ledger", "cash", "mean_entry", "borrowed", "risk", "done", "trans_units", "price"):
            assert torch.equal(envs[0].t[name], envs[1].t[name]), (t, name)
        seen |= set(torch.unique(envs[0].t["risk"]).tolist())
    assert A.RISK_INSUFF_MARGIN in seen and A.RISK_MARGIN_CALL in seen, seen


def test_single_step_full_size_vs_oracle():
    """65,536 x 16 (the bench launch size): two steps compared with the oracle element for element -- grid tail,
    indexing and the block-level carry chain at full size."""
    import os
    from madigan_b200.environments import Env
    from oracle.oracle import OracleBatch
    pairs = {f"pair{i}": {"data_source_type": "OUPair", "data_source_config": {"theta": .015, "phi": .01, "noise": .03}}
             for i in range(8)}
    N, nA = 65_536 + 7, 16  # a ragged last tile
    P, _ = make_params("Composite", pairs, required_margin=.5, maintenance_margin=.25, transaction_cost_rel=.02,
                       slippage_rel=.001)
    rw = dict(reward_shaper_config={"reward_shaper": "DSR", "adaptation_rate": .001}, nstep_return=1, reduce_rewards=True)
    from madigan_b200.environments.data_source import make_reward
    R = make_reward(rw["reward_shaper_config"], 1, .99, True, n_assets=nA)
    env = Env("Composite", INIT_CASH, {"data_source_config": pairs}, n_envs=N, window=4, seed=3, reward=rw)
    env.setRequiredMargin(.5); env.setMaintenanceMargin(.25); env.setTransactionCost(.02, 0.); env.setSlippage(.001, 0.)
    orc = OracleBatch(N, P, R, window=4, seed=3, threads=os.cpu_count() or 1)
    orc.reset(fill_ticks=1, clear_nstep=False)
    st = orc.state()
    for name in ("price", "gstate", "timestamp"):
        env.t[name].copy_(torch.from_numpy(st[name]))
    env.invalidate()
    rng = np.random.default_rng(1)
    for t in range(2):
        units = rng.integers(-1, 2, size=(N, nA)).astype(np.float64) * 30_000. * rng.uniform(.2, 1.5, size=(N, nA))
        z = rng.standard_normal((P.n_normals, N))
        env.step(torch.from_numpy(units), normals=z, uniforms=np.zeros((1, N)))
        orc.step(units, normals=z, uniforms=np.zeros((1, N)))
        s2 = orc.state()
        for name in ("ledger", "mean_entry", "borrowed", "cash", "price"):
            assert np.array_equal(cpu(env.t[name]), s2[name]), f"step {t}: {name}"
        assert np.array_equal(cpu(env.t["risk"]), orc.risk) and np.array_equal(cpu(env.t["done"]), orc.done)
        assert np.array_equal(cpu(env.t["trans_units"]), orc.trans_units)
        np.testing.assert_allclose(cpu(env.t["reward"]), orc.reward, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(cpu(env.t["shaped_reward"]), orc.shaped_reward, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(cpu(env.t["obs_port"][env.head]), orc.obs_port[orc.head], rtol=1e-12, atol=1e-14)
