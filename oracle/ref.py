"""TEST INFRASTRUCTURE, NOT PRODUCT CODE: ctypes front-end of oracle/_ref/libmadigan_ref.so -- the REFERENCE's own
C++ env (Env / Broker / Account / Portfolio / DataSource sources compiled unmodified, see oracle/ref_build.py).
Only tests/, __graft_entry__ and bench.py's baseline legs may import this."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libmadigan_ref.so")
_LIB = None

SYNTH, OU, OUPAIR, MULTIPAIR, SIMPLETREND, TRENDOU, TRENDYOU, SAWTOOTH, TRIANGLE, GAUSSIAN = range(10)
SINEADDER, SINEDYNAMIC, SINEDYNAMICTREND = 10, 11, 12  # one asset; n_assets = number of components


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        vp, dp, ip = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.ref_env_create.restype = vp
        L.ref_env_create.argtypes = [C.c_int, C.c_int, dp, C.c_double, C.c_uint]
        L.ref_env_destroy.argtypes = [vp]
        L.ref_env_set.argtypes = [vp] + [C.c_double] * 6
        L.ref_env_next_normals.restype = C.c_int
        L.ref_env_next_normals.argtypes = [vp, dp]
        L.ref_env_next_draws.restype = C.c_int
        L.ref_env_next_draws.argtypes = [vp, dp, dp, C.c_int]
        L.ref_env_trend_state.argtypes = [vp, dp, ip, ip, ip]
        L.ref_env_next_sine_draws.restype = C.c_int
        L.ref_env_next_sine_draws.argtypes = [vp, dp, dp, dp, C.c_int]
        L.ref_env_sine_state.argtypes = [vp, dp, dp, ip, ip, ip]
        L.ref_env_reset.argtypes = [vp, dp, dp, C.POINTER(C.c_longlong)]
        L.ref_env_step.argtypes = [vp, C.c_int, dp, C.c_int, dp, dp, C.POINTER(C.c_longlong), dp, ip, dp, dp, dp, ip, ip]
        L.ref_env_accounting.argtypes = [vp, dp, dp, dp]
        L.ref_env_run.restype = C.c_longlong
        L.ref_env_run.argtypes = [vp, C.c_longlong, dp, C.c_int]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefEnv:
    """One reference Env (madigan/environments/cpp/Env.h) with a re-seeded generator."""

    def __init__(self, kind, n_assets, params, init_cash=1_000_000., seed=12345):
        self.L = lib()
        p = np.ascontiguousarray(params, dtype=np.float64)
        self.n = 2 if kind == OUPAIR else (1 if kind >= SINEADDER else int(n_assets))
        self.ncomp = int(n_assets) if kind >= SINEADDER else 0
        self.h = self.L.ref_env_create(kind, int(n_assets), _dp(p), float(init_cash), int(seed))

    def __del__(self):
        try:
            self.L.ref_env_destroy(self.h)
        except Exception:
            pass

    def set(self, reqM, maintM, tc_rel=0., tc_abs=0., sl_rel=0., sl_abs=0.):
        self.L.ref_env_set(self.h, reqM, maintM, tc_rel, tc_abs, sl_rel, sl_abs)

    def next_normals(self):
        z = np.zeros(64)
        n = self.L.ref_env_next_normals(self.h, _dp(z))
        return z[:n].copy()

    def next_draws(self, after_reset=False):
        """trend sources: (normals[n], uniforms[4n]) of the next getData(), in the oracle's slot order"""
        z, u = np.zeros(self.n), np.zeros(4 * self.n)
        n = self.L.ref_env_next_draws(self.h, _dp(z), _dp(u), int(after_reset))
        assert n == self.n
        return z, u

    def next_sine_draws(self, after_reset=False):
        """SineAdder / SineDynamic(Trend): (normals, uniforms, ctor_uniforms) of the next getData() -- after a reset()
        first when after_reset -- in the oracle's slot order (see ref_driver.cpp)"""
        z, u, cu = np.zeros(max(self.ncomp, 1)), np.ones(16), np.zeros(3 * self.ncomp)
        n = self.L.ref_env_next_sine_draws(self.h, _dp(z), _dp(u), _dp(cu), int(after_reset))
        assert n >= 1
        return z[:n].copy(), u, cu

    def sine_state(self, n_trends=0):
        ip = C.POINTER(C.c_int)
        comp, tc = np.zeros(4 * self.ncomp), np.zeros(1)
        d, ln, tr = (np.zeros(max(n_trends, 1), dtype=np.int32) for _ in range(3))
        self.L.ref_env_sine_state(self.h, _dp(comp), _dp(tc), d.ctypes.data_as(ip), ln.ctypes.data_as(ip), tr.ctypes.data_as(ip))
        return dict(comp=comp, trend_component=tc[0], direction=d[:n_trends], length=ln[:n_trends],
                    trending=tr[:n_trends].astype(bool))

    def trend_state(self):
        ip = C.POINTER(C.c_int)
        dY = np.zeros(self.n)
        d, ln, tr = (np.zeros(self.n, dtype=np.int32) for _ in range(3))
        self.L.ref_env_trend_state(self.h, _dp(dY), d.ctypes.data_as(ip), ln.ctypes.data_as(ip), tr.ctypes.data_as(ip))
        return dict(dY=dY, direction=d, length=ln, trending=tr.astype(bool))

    def reset(self):
        price, port, ts = np.zeros(self.n), np.zeros(self.n + 1), C.c_longlong()
        self.L.ref_env_reset(self.h, _dp(price), _dp(port), C.byref(ts))
        return dict(price=price, portfolio=port, timestamp=ts.value)

    def step(self, units=None, asset_idx=None):
        n = self.n
        price, port, ts = np.zeros(n), np.zeros(n + 1), C.c_longlong()
        reward, done, mc = C.c_double(), C.c_int(), C.c_int()
        tp, tu, tc = np.zeros(n), np.zeros(n), np.zeros(n)
        risk = np.zeros(n, dtype=np.int32)
        if units is None:
            mode, u, ai = 0, np.zeros(1), 0
        elif asset_idx is not None:
            mode, u, ai = 2, np.array([units], dtype=np.float64), int(asset_idx)
        else:
            mode, u, ai = 1, np.ascontiguousarray(units, dtype=np.float64), 0
        self.L.ref_env_step(self.h, mode, _dp(u), ai, _dp(price), _dp(port), C.byref(ts), C.byref(reward),
                            C.byref(done), _dp(tp), _dp(tu), _dp(tc), risk.ctypes.data_as(C.POINTER(C.c_int)),
                            C.byref(mc))
        return dict(price=price, portfolio=port, timestamp=ts.value, reward=reward.value, done=bool(done.value),
                    transactionPrice=tp, transactionUnits=tu, transactionCost=tc, riskInfo=risk,
                    marginCall=bool(mc.value))

    def accounting(self):
        out, led, mep = np.zeros(8), np.zeros(self.n), np.zeros(self.n)
        self.L.ref_env_accounting(self.h, _dp(out), _dp(led), _dp(mep))
        keys = ("equity", "cash", "pnl", "balance", "availableMargin", "usedMargin", "borrowedMargin",
                "borrowedAssetValue")
        d = dict(zip(keys, out))
        d["ledger"], d["meanEntryPrices"] = led, mep
        return d

    def run(self, steps, acts):
        """`steps` calls of Env::step(units) cycling over acts (n_act, nA), reset on done; returns #resets."""
        a = np.ascontiguousarray(acts, dtype=np.float64)
        return int(self.L.ref_env_run(self.h, int(steps), _dp(a), a.shape[0]))
