"""The reference's own ledger known-answer tests, replayed against the CPU oracle.

Every test restates one test of the reference's
madigan/environments/cpp/tests/envTest.py (line ranges cited per test) or
envTest.cpp, with the oracle's Portfolio/Broker restatement in place of the pybind
classes.  These pin the oracle's stage-2 arithmetic before
anything is compared against it.
"""
import numpy as np
import pytest
from numpy.testing import assert_allclose

from madigan_b200 import _abi as A
from madigan_b200.environments.data_source import make_params
from oracle.oracle import OracleEnv

SYNTH_CFG = {'freq': [1., 0.3, 2., 0.5], 'mu': [2., 2.1, 2.2, 2.3], 'amp': [1., 1.2, 1.3, 1.0],
             'phase': [0., 1.0, 2., 1.], 'dX': 0.01, 'noise': 0.}  # envTest.py:11-21


def portfolio(required_margin=1., maintenance_margin=.25, init_cash=1_000_000, cfg=None, tick=True):
    """Portfolio(assets, initCash) + setDataSource(Synth()) ; Portfolio defaults Portfolio.h:121-122."""
    P, _ = make_params("Synth", cfg, init_cash=init_cash, required_margin=required_margin,
                       maintenance_margin=maintenance_margin)
    return OracleEnv(P, construct=tick)


def ref_transaction(units, init_cash, prices, assetIdx=0, margin=1.):
    """envTest.py:102-115, verbatim arithmetic."""
    cash = init_cash
    price = prices[assetIdx]
    cost = margin * (price * units)
    cash -= cost
    borrowed_margin = (1 - margin) * (price * units)
    if borrowed_margin < 0.:
        cash -= borrowed_margin
        borrowed_margin = 0.
    equity = cash + units * (price) - borrowed_margin
    return cash, borrowed_margin, equity


def test_datasource_default_equals_config():
    """envTest.py:33-48: default ctor == explicit params == dict config."""
    p1 = portfolio(cfg=None).prices.copy()
    p2 = portfolio(cfg=SYNTH_CFG).prices.copy()
    np.testing.assert_equal(p1, p2)
    # and the values are the noise-free sines of DataSource.cpp:535-543 at x = phase
    x = np.array(SYNTH_CFG['phase'])
    exp = np.array(SYNTH_CFG['mu']) + np.array(SYNTH_CFG['amp']) * np.sin(
        (3.141592653589793238463 * 2) * x * np.array(SYNTH_CFG['freq']))
    assert_allclose(p1, exp, rtol=1e-15)


@pytest.mark.parametrize("units,margin", [(1000., 1.), (-1000., 1.), (1000., .1), (-1000., .1)])
def test_port_accounting_logic(units, margin):
    """envTest.py:119-145 (and :255-281, :303-330 which route the same call through Account/Broker)."""
    port = portfolio(required_margin=margin, cfg=SYNTH_CFG)
    prices = port.prices.copy()
    port.handleTransaction(0, prices[0], units, 0.)
    cash, bm, eq = ref_transaction(units, 1_000_000, prices, 0, margin)
    assert cash == port.cash
    assert bm == port.borrowedMargin
    assert eq == port.equity


@pytest.mark.parametrize("units,margin", [(1000., 1.), (-1000., 1.), (1000., .1), (-1000., .1)])
def test_broker_accounting_logic(units, margin):
    """envTest.py:303-330: Broker.handleTransaction(assetIdx, units), zero cost/slippage."""
    port = portfolio(required_margin=margin, cfg=SYNTH_CFG)
    prices = port.prices.copy()
    resp = port.brokerTransaction(0, units)
    cash, bm, eq = ref_transaction(units, 1_000_000, prices, 0, margin)
    assert resp["riskInfo"] == A.RISK_GREEN
    assert cash == port.cash
    assert bm == port.borrowedMargin
    assert eq == port.equity


def test_port_ledger():
    """envTest.py:148-178."""
    ATOL = 1e-8
    port1 = portfolio(required_margin=1.)
    prices = port1.prices.copy()
    for i, u in zip([0, 1, 2, 3], [1000, 2000, -4000, 1000]):
        port1.handleTransaction(i, prices[i], u, 0.)
    assert abs((1 - port1.ledgerNormed.sum()) * port1.equity - (port1.cash - port1.borrowedMargin)) < ATOL
    assert abs((1 - port1.ledgerNormed.sum()) * port1.equity - port1.cash) < ATOL
    assert abs(port1.ledgerNormedFull.sum() - 1.) < ATOL
    port2 = portfolio(required_margin=.1)
    for i, u in zip([0, 1, 2, 3], [1000, 2000, -4000, 1000]):
        port2.handleTransaction(i, prices[i], u, 0.)
    assert abs((1 - port2.ledgerNormed.sum()) * port2.equity - (port2.cash - port2.borrowedMargin)) < ATOL
    renormed = port1.ledgerAbsNormedFull * 1 / port1.ledgerAbsNormedFull.sum()
    assert_allclose(renormed, port1.ledgerNormedFull)


def test_successive_accounting1():
    """envTest.py:404-440: buy 10k,10k, sell 20k, sell 20k, buy 10k,10k."""
    port = portfolio(required_margin=0.1)
    prices = port.prices.copy()
    port.handleTransaction(0, prices[0], 10_000)
    assert port.assetValue == prices[0] * 10_000
    port.handleTransaction(0, prices[0], 10_000)
    assert port.cash == 1_000_000. - 0.1 * prices[0] * 20_000
    assert port.assetValue == prices[0] * 20_000
    assert port.usedMargin == 0.1 * prices[0] * 20_000
    assert port.borrowedMargin == 0.9 * prices[0] * 20_000
    assert port.borrowedAssetValue == 0.
    port.handleTransaction(0, prices[0], -20_000)
    assert_allclose(port.cash, 1_000_000., rtol=1e-12)
    assert_allclose(port.assetValue, 0., rtol=1e-12)
    assert_allclose(port.usedMargin, 0., rtol=1e-12)
    assert_allclose(port.borrowedMargin, 0., rtol=1e-12)
    assert_allclose(port.borrowedAssetValue, 0., rtol=1e-12)
    port.handleTransaction(0, prices[0], -20_000)
    assert_allclose(port.cash, 1_000_000 + prices[0] * 20_000., rtol=1e-12)
    assert_allclose(port.assetValue, prices[0] * -20_000, rtol=1e-12)
    assert_allclose(port.usedMargin, 0.1 * prices[0] * 20_000, rtol=1e-12)
    assert_allclose(port.borrowedMargin, 0.)
    assert_allclose(port.borrowedAssetValue, prices[0] * -20_000, rtol=1e-12)
    port.handleTransaction(0, prices[0], 10_000)
    port.handleTransaction(0, prices[0], 10_000)
    assert_allclose(port.cash, 1_000_000., rtol=1e-12)
    assert_allclose(port.assetValue, 0., rtol=1e-12)
    assert_allclose(port.usedMargin, 0., rtol=1e-12)
    assert_allclose(port.borrowedMargin, 0., rtol=1e-12)
    assert_allclose(port.borrowedAssetValue, 0., rtol=1e-12)


def test_successive_accounting2():
    """envTest.py:443-456: short 10k then buy 20k -> reversed to +10k."""
    port = portfolio(required_margin=0.1)
    prices = port.prices.copy()
    port.handleTransaction(0, prices[0], -10_000)
    port.handleTransaction(0, prices[0], 20_000)
    assert port.cash == 1_000_000. - 0.1 * prices[0] * 10_000
    assert port.assetValue == prices[0] * 10_000
    assert port.usedMargin == 0.1 * prices[0] * 10_000
    assert port.borrowedMargin == 0.9 * prices[0] * 10_000
    assert port.borrowedAssetValue == 0.


def test_successive_accounting3():
    """envTest.py:459-472: long 10k then sell 20k -> reversed to -10k."""
    port = portfolio(required_margin=0.1)
    prices = port.prices.copy()
    port.handleTransaction(0, prices[0], 10_000)
    port.handleTransaction(0, prices[0], -20_000)
    assert port.cash == 1_000_000. + prices[0] * 10_000
    assert port.assetValue == -prices[0] * 10_000
    assert port.usedMargin == 0.1 * prices[0] * 10_000
    assert port.borrowedMargin == 0.
    assert port.borrowedAssetValue == -prices[0] * 10_000


def test_multiasset_accounting():
    """envTest.py:475-509."""
    port = portfolio(required_margin=0.1)
    prices = port.prices.copy()
    port.handleTransaction(0, prices[0], 20_000)
    assert port.cash == 1_000_000. - 0.1 * prices[0] * 20_000
    assert port.assetValue == prices[0] * 20_000
    assert port.usedMargin == 0.1 * prices[0] * 20_000
    assert port.borrowedMargin == 0.9 * prices[0] * 20_000
    assert port.borrowedAssetValue == 0.
    port.handleTransaction(3, prices[3], -20_000)
    cash = 1_000_000. - (0.1 * prices[0] * 20_000) + (prices[3] * 20_000)
    balance = 1_000_000 - (0.1 * prices[0] * 20_000)
    assetValue = prices[0] * 20_000 + prices[3] * -20_000
    usedMargin = 0.1 * prices[0] * 20_000 + 0.1 * prices[3] * 20_000
    assert_allclose(port.cash, cash, rtol=1e-12)
    assert_allclose(port.balance, balance, rtol=1e-12)
    assert_allclose(port.assetValue, assetValue, rtol=1e-12)
    assert_allclose(port.usedMargin, usedMargin, rtol=1e-12)
    assert_allclose(port.borrowedMargin, 0.9 * prices[0] * 20_000, rtol=1e-12)
    assert_allclose(port.borrowedAssetValue, prices[3] * -20_000, rtol=1e-12)


def test_port_risk_handling():
    """envTest.py:512-547: margin-call thresholds at price 4 -> 3.71 / 3.69."""
    port = portfolio(required_margin=0.1, maintenance_margin=1.)
    prices = port.prices  # live view into the oracle's buffer, like the pybind reference (envTest.py:516)
    prices[1] = 4
    price = prices[1]
    reqM = 0.1
    port.handleTransaction(1, price, 1_000_000)
    assert port.checkRisk(1, (-1. + (port.balance + port.pnl) / reqM) / price) == A.RISK_GREEN
    assert port.checkRisk(1, (0. + (port.balance + port.pnl) / reqM) / price) == A.RISK_INSUFF_MARGIN
    assert port.checkRisk(1, (1. + (port.balance + port.pnl) / reqM) / price) == A.RISK_INSUFF_MARGIN
    new_price = 3.71
    prices[1] = new_price
    assert port.checkRisk() == A.RISK_GREEN
    assert port.checkRisk(1, 1_000_000 / price) == A.RISK_GREEN
    new_price = 3.69
    prices[1] = new_price
    assert port.checkRisk() == A.RISK_MARGIN_CALL
    assert port.checkRisk(1, (-1. + (port.balance + port.pnl) / reqM) / price) == A.RISK_MARGIN_CALL
    assert port.checkRisk(1, 0.) == A.RISK_MARGIN_CALL
    loss = 1_000_000 * (price - new_price)
    equity = 1_000_000 - loss
    assert_allclose(-loss, port.pnl, rtol=1e-12)
    port.handleTransaction(1, new_price, -1_000_000)
    assert_allclose(equity, port.equity, rtol=1e-12)
    assert_allclose(equity, port.cash, rtol=1e-12)


def test_broker_risk_handling():
    """envTest.py:550-566: response fields of a green transaction."""
    port = portfolio(required_margin=0.1, maintenance_margin=1.)
    port.prices[1] = 4
    resp = port.brokerTransaction(1, 1_000_000)
    assert resp["transactionPrice"] == 4
    assert resp["transactionCost"] == 0.
    assert resp["riskInfo"] == A.RISK_GREEN


def test_cpp_accounting_identities():
    """envTest.cpp:212-266: pnl==0 at entry, pnl tracks price, three equity identities < 1e-10."""
    port = portfolio(required_margin=0.1)
    p0 = port.prices[0]
    port.handleTransaction(0, p0, 10_000, 0.)
    assert port.pnl == 0.
    port.tick()
    p1 = port.prices[0]
    assert abs(port.pnl - 10_000 * (p1 - p0)) < 1e-7
    eq = port.equity
    assert abs(eq - (port.cash + port.assetValue - port.borrowedMargin)) < 1e-10
    assert abs(eq - (port.balance + port.pnl + port.usedMargin)) < 1e-10
    # add to the position, then reverse it: pnl of a fresh entry is zero again
    port.handleTransaction(0, p1, -20_000, 0.)
    assert abs(port.pnl) < 1e-9
    eq = port.equity
    assert abs(eq - (port.cash + port.assetValue - port.borrowedMargin)) < 1e-10


def test_env_smoke():
    """envTest.py:579-599: Env.step(), step(units), step(i, units) run and keep the identities."""
    P, _ = make_params("Synth", SYNTH_CFG, required_margin=1., maintenance_margin=.25)
    env = OracleEnv(P)
    env.step()
    o = env.step([10000, 20000, -20000, -40000])
    assert o["riskInfo"].tolist() == [0, 0, 0, 0]
    assert abs(o["portfolio"].sum() - 1.) < 1e-12
    env.step(10_000, asset_idx=0)
    env.step(-20_000, asset_idx=0)
    env2 = OracleEnv(P)
    o = env2.step(-1 + 1_000_000 / env2.prices[0], asset_idx=0)
    assert o["riskInfo"][0] == A.RISK_GREEN
