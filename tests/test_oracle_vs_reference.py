"""The CPU oracle against the REFERENCE'S OWN C++ env, compiled unmodified from its sources
(oracle/ref_build.py -> oracle/_ref/libmadigan_ref.so: Env.h, Broker.cpp, Account.cpp, Portfolio.cpp,
DataSource.cpp, Config.cpp against a small Eigen shim with left-to-right folds, strict IEEE flags).

Both sides get the same actions; the reference's generator is re-seeded and a clone of its engine +
distribution objects yields the standard normals it is about to draw, which are injected into the oracle.
Every output of Env::step -- prices, ledgerNormedFull, reward, done, transaction price/units/cost, risk codes,
marginCall -- and the accounting properties must then agree BIT FOR BIT (same compiler, same libm).
This pins what the reference's own tests leave unpinned: OU / OUPair / noisy Synth generators, Env::step's
reward clamp and done logic, slippage and transaction-cost arithmetic, the risk gates under leverage.

Skipped when the reference build is absent (it can only be produced where /root/reference exists)."""
import numpy as np
import pytest

from madigan_b200 import _abi as A
from madigan_b200.environments.data_source import make_params
from oracle import ref
from oracle.oracle import OracleEnv

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (needs /root/reference)")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.array_equal(bits(a), bits(b)) or np.array_equal(a, b, equal_nan=True)


CASES = {
    "synth": (ref.SYNTH, "Synth", dict(freq=[1., 0.3, 2.], mu=[2., 2.1, 2.2], amp=[1., 1.2, 1.3], phase=[0., 1., 2.],
                                       dX=0.01, noise=0.05)),
    "ou": (ref.OU, "OU", dict(mean=[10., 5., 1., 2.], theta=[.08, .15, .15, .1], phi=[.04, .02, .01, .03])),
    "oupair": (ref.OUPAIR, "OUPair", dict(theta=.015, phi=.01, noise=.03)),
    "pairs8": (ref.MULTIPAIR, "Composite", dict(theta=.015, phi=.01, noise=.03)),
}


def build(case, margins, costs, seed):
    kind, ds_type, cfg = CASES[case]
    if kind == ref.SYNTH:
        n = len(cfg["freq"])
        p = [x for i in range(n) for x in (cfg["freq"][i], cfg["mu"][i], cfg["amp"][i], cfg["phase"][i])]
        p += [cfg["dX"], cfg["noise"]]
        ds_cfg = cfg
    elif kind == ref.OU:
        n = len(cfg["mean"])
        p = [x for i in range(n) for x in (cfg["mean"][i], cfg["theta"][i], cfg["phi"][i])]
        ds_cfg = cfg
    elif kind == ref.OUPAIR:
        n, p, ds_cfg = 2, [cfg["theta"], cfg["phi"], cfg["noise"]], cfg
    else:
        n, p = 16, [cfg["theta"], cfg["phi"], cfg["noise"]]
        ds_cfg = {f"pair{i}": {"data_source_type": "OUPair", "data_source_config": cfg} for i in range(8)}
    r = ref.RefEnv(kind, n, p, 1_000_000., seed)
    r.set(margins[0], margins[1], *costs)
    P, _ = make_params(ds_type, ds_cfg, required_margin=margins[0], maintenance_margin=margins[1],
                       transaction_cost_rel=costs[0], transaction_cost_abs=costs[1], slippage_rel=costs[2],
                       slippage_abs=costs[3])
    o = OracleEnv(P, construct=False)
    return r, o, P.n_assets


def gen_units(rng, o, n, scale):
    price = np.where(np.abs(o.prices) > 1e-9, o.prices, 1.0)
    led = o.ledger.copy()
    a = rng.integers(-1, 2, size=n).astype(np.float64)
    units = a * (scale / np.abs(price)) * rng.uniform(0.2, 1.5, size=n)
    r = rng.random(n)
    units = np.where(r < 0.08, -led, units)
    units = np.where((r >= 0.08) & (r < 0.14), -2. * led, units)
    units = np.where((r >= 0.14) & (r < 0.17), -0.5 * led, units)
    units = np.where((r >= 0.17) & (r < 0.20), units * 50., units)
    return units


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("margins,costs,scale", [
    ((1., .25), (0., 0., 0., 0.), 30_000.),
    ((.1, .25), (.02, 1.5, .001, .002), 400_000.),
    ((.05, 1.), (.001, 0., 0., 0.), 900_000.),
])
def test_env_step_bit_exact_vs_reference(case, margins, costs, scale):
    rng = np.random.default_rng(hash((case, margins[0])) % 2 ** 31)
    r, o, n = build(case, margins, costs, seed=4242)
    z = r.next_normals()
    rs, os_ = r.reset(), o.reset(normals=z)
    assert same(rs["price"], os_["price"]) and same(rs["portfolio"], os_["portfolio"])
    seen, dones = set(), 0
    for t in range(400):
        mode = t % 7
        z = r.next_normals()
        if mode == 5:
            ro, oo = r.step(), o.step(normals=z)
        elif mode == 6:
            i = int(rng.integers(0, n))
            u = float(gen_units(rng, o, n, scale)[i])
            ro, oo = r.step(u, asset_idx=i), o.step(u, asset_idx=i, normals=z)
        else:
            u = gen_units(rng, o, n, scale)
            ro, oo = r.step(u), o.step(u, normals=z)
        for k in ("price", "portfolio", "transactionPrice", "transactionUnits", "transactionCost"):
            assert same(ro[k], oo[k]), f"step {t} {k}"
        assert same(ro["reward"], oo["reward"]), f"step {t} reward {ro['reward']} {oo['reward']}"
        assert ro["done"] == oo["done"], f"step {t} done"
        assert np.array_equal(ro["riskInfo"], oo["riskInfo"]), f"step {t} riskInfo"
        if mode != 5:
            assert ro["marginCall"] == oo["marginCall"], f"step {t} marginCall"
        acc = r.accounting()
        assert same(acc["ledger"], o.ledger) and same(acc["meanEntryPrices"], o.meanEntryPrices)
        for k, v in (("equity", o.equity), ("cash", o.cash), ("pnl", o.pnl), ("balance", o.balance),
                     ("availableMargin", o.availableMargin), ("usedMargin", o.usedMargin),
                     ("borrowedMargin", o.borrowedMargin), ("borrowedAssetValue", o.borrowedAssetValue)):
            assert same(acc[k], v), f"step {t} {k}: {acc[k]!r} vs {v!r}"
        seen |= set(int(x) for x in oo["riskInfo"])
        if oo["done"]:
            dones += 1
            z = r.next_normals()
            rs, os_ = r.reset(), o.reset(normals=z)
            assert same(rs["price"], os_["price"]) and same(rs["portfolio"], os_["portfolio"])
    if margins[0] < 1.:
        assert A.RISK_INSUFF_MARGIN in seen


def test_reference_timestamp_counts_ticks():
    """The reference leaves timestamp_ uninitialised (quirk A8); its increments are still one per getData."""
    r, o, n = build("oupair", (1., .25), (0., 0., 0., 0.), seed=1)
    t0 = r.reset()["timestamp"]
    t1 = r.step()["timestamp"]
    t2 = r.step(np.zeros(2))["timestamp"]
    assert t1 - t0 == 1 and t2 - t1 == 1
