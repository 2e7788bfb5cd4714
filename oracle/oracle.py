"""TEST INFRASTRUCTURE, NOT PRODUCT CODE: ctypes front-end of oracle/mdg_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  It drives the plain-C restatement of the reference's
Env.step path (one env at a time, fp64, injected noise) with numpy arrays laid out
exactly like the CUDA path's tensors ([rows][N]).
"""
import ctypes as C
import os
import subprocess

import numpy as np

from madigan_b200 import _abi as A

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ORC_MAX_GSTATE = 320
ORC_MAX_CTOR_U = 3 * A.MDG_MAX_SINE_COMPONENTS


class OrcEnv(C.Structure):
    _fields_ = [("P", A.MdgParams), ("R", A.MdgReward),
                ("price", C.c_double * A.MDG_MAX_ASSETS), ("ledger", C.c_double * A.MDG_MAX_ASSETS),
                ("mep", C.c_double * A.MDG_MAX_ASSETS), ("bm", C.c_double * A.MDG_MAX_ASSETS),
                ("cash", C.c_double), ("gstate", C.c_double * ORC_MAX_GSTATE),
                ("timestamp", C.c_int64), ("seed", C.c_uint64), ("gid", C.c_int64),
                ("A", C.c_double * A.MDG_MAX_ASSETS), ("B", C.c_double * A.MDG_MAX_ASSETS),
                ("ring", (C.c_double * A.MDG_MAX_ASSETS) * A.MDG_MAX_NSTEP),
                ("ring_len", C.c_int32), ("n_ctor_u", C.c_int32), ("ctor_u", C.c_double * ORC_MAX_CTOR_U)]


def host_params(params):
    """The oracle reads MdgParams.gen_ext on the host: a copy of `params` whose gen_ext points at the numpy
    table make_params attached (``ext_host``).  Returns (params, keepalive)."""
    ext = getattr(params, "ext_host", None)
    if ext is None:
        return params, None
    P = A.MdgParams.from_buffer_copy(params)
    ext = np.ascontiguousarray(ext, dtype=np.float64)
    P.gen_ext = ext.ctypes.data_as(C.c_void_p)
    P.n_gen_ext = ext.size
    P.ext_host = ext
    return P, ext


class OrcStepOut(C.Structure):
    _fields_ = [("price", C.c_double * A.MDG_MAX_ASSETS), ("port", C.c_double * (A.MDG_MAX_ASSETS + 1)),
                ("timestamp", C.c_int64), ("reward", C.c_double), ("done", C.c_uint8),
                ("tp", C.c_double * A.MDG_MAX_ASSETS), ("tu", C.c_double * A.MDG_MAX_ASSETS),
                ("tc", C.c_double * A.MDG_MAX_ASSETS), ("risk", C.c_uint8 * A.MDG_MAX_ASSETS),
                ("margin_call", C.c_uint8), ("agent_reward", C.c_double * A.MDG_MAX_ASSETS),
                ("shaped", (C.c_double * A.MDG_MAX_ASSETS) * A.MDG_MAX_NSTEP),
                ("n_popped", C.c_int32)]


def build(force=False):
    """Compile oracle/libmdg_oracle.so (gcc, -ffp-contract=off)."""
    so = os.path.join(_HERE, "libmdg_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("mdg_oracle.c", "mdg_oracle.h", "reward_norm_oracle.c")]
    src.append(os.path.join(_HERE, "..", "include", "madigan_b200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.run(["make", "-s", "-C", _HERE], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dbl = C.c_double
        pe = C.POINTER(OrcEnv)
        for name in ("orc_equity", "orc_asset_value", "orc_pnl", "orc_balance", "orc_available_margin",
                     "orc_used_margin", "orc_borrowed_margin", "orc_borrowed_asset_value"):
            getattr(L, name).restype = dbl
            getattr(L, name).argtypes = [pe]
        for name in ("orc_ledger_normed", "orc_ledger_normed_full", "orc_ledger_abs_normed",
                     "orc_ledger_abs_normed_full"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [pe, C.POINTER(dbl)]
        L.orc_init.argtypes = [pe, C.POINTER(A.MdgParams), C.POINTER(A.MdgReward), C.c_uint64, C.c_int64]
        L.orc_init.restype = None
        L.orc_init_inject.argtypes = [pe, C.POINTER(A.MdgParams), C.POINTER(A.MdgReward), C.c_uint64, C.c_int64,
                                      C.c_void_p, C.c_int]
        L.orc_init_inject.restype = None
        L.orc_set_ctor_uniforms.argtypes = [pe, C.c_void_p, C.c_int]
        L.orc_set_ctor_uniforms.restype = None
        L.orc_tick.argtypes = [pe, C.c_void_p, C.c_void_p]
        L.orc_tick.restype = None
        L.orc_reset.argtypes = [pe, C.c_void_p, C.c_void_p, C.POINTER(OrcStepOut)]
        L.orc_reset.restype = None
        L.orc_step.argtypes = [pe, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(OrcStepOut)]
        L.orc_step.restype = None
        L.orc_shaper_feed.argtypes = [pe, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(OrcStepOut)]
        L.orc_shaper_feed.restype = None
        L.orc_check_risk.argtypes = [pe]
        L.orc_check_risk.restype = C.c_int
        L.orc_check_risk_asset.argtypes = [pe, C.c_int, dbl]
        L.orc_check_risk_asset.restype = C.c_int
        L.orc_handle_transaction.argtypes = [pe, C.c_int, dbl, dbl, dbl]
        L.orc_handle_transaction.restype = None
        L.orc_broker_transaction.argtypes = [pe, C.c_int, dbl, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl)]
        L.orc_broker_transaction.restype = C.c_int
        L.orc_philox4x32_10.argtypes = [C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.orc_philox4x32_10.restype = None
        L.orc_draw_normal.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int]
        L.orc_draw_normal.restype = dbl
        L.orc_draw_uniform.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int]
        L.orc_draw_uniform.restype = dbl
        L.orc_action_units.argtypes = [pe, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.orc_weight_units.argtypes = [pe, C.c_void_p, C.c_void_p]
        L.orc_weight_units.restype = None
        L.orc_batch_create.argtypes = [C.c_int64, C.POINTER(A.MdgParams), C.POINTER(A.MdgReward), C.c_uint64, C.c_int64]
        L.orc_batch_create.restype = C.c_void_p
        L.orc_batch_destroy.argtypes = [C.c_void_p]
        L.orc_batch_destroy.restype = None
        L.orc_batch_env.argtypes = [C.c_void_p, C.c_int64]
        L.orc_batch_env.restype = pe
        L.orc_batch_export_state.argtypes = [C.c_void_p, C.POINTER(A.MdgState)]
        L.orc_batch_export_state.restype = None
        L.orc_batch_step.argtypes = [C.c_void_p, C.POINTER(A.MdgStepIO), C.POINTER(A.MdgLaunch), C.c_int]
        L.orc_batch_step.restype = None
        L.orc_batch_reset.argtypes = [C.c_void_p, C.POINTER(A.MdgStepIO), C.POINTER(A.MdgLaunch), C.c_void_p,
                                      C.c_int, C.c_int, C.c_int]
        L.orc_batch_reset.restype = None
        L.orc_batch_derived.argtypes = [C.c_void_p, C.POINTER(A.MdgDerived)]
        L.orc_batch_derived.restype = None
        L.orc_reward_norm_reset.argtypes = [C.POINTER(A.MdgRewardNorm), C.c_void_p]
        L.orc_reward_norm_reset.restype = None
        L.orc_reward_norm_stream.argtypes = [C.POINTER(A.MdgRewardNorm), C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_reward_norm_stream.restype = None
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleEnv:
    """One env: the reference's Env / Portfolio / Broker, restated.  Mirrors the members the
    reference's own tests touch (environments/cpp/tests/envTest.py)."""

    def __init__(self, params, reward=None, seed=0, gid=0, construct=True, ctor_uniforms=None):
        self.L = lib()
        self.e = OrcEnv()
        params, self._ext = host_params(params)
        self.P = params
        cu = None if ctor_uniforms is None else np.ascontiguousarray(ctor_uniforms, dtype=np.float64)
        self.L.orc_init_inject(C.byref(self.e), C.byref(params), C.byref(reward) if reward is not None else None,
                               seed, gid, _ptr(cu), 0 if cu is None else cu.size)
        self.nA = params.n_assets
        if construct:  # Env ctor consumes one tick (Env.h:160)
            self.tick()

    # -- setters (Env.h:94-111)
    def setRequiredMargin(self, v): self.e.P.required_margin = v
    def setMaintenanceMargin(self, v): self.e.P.maintenance_margin = v
    def setTransactionCost(self, rel=0., ab=0.): self.e.P.tcost_rel, self.e.P.tcost_abs = rel, ab
    def setSlippage(self, rel=0., ab=0.): self.e.P.slippage_rel, self.e.P.slippage_abs = rel, ab

    def tick(self, normals=None, uniforms=None):
        n = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        self.L.orc_tick(C.byref(self.e), _ptr(n), _ptr(u))
        return self.prices

    def _out(self, o):
        nA = self.nA
        return dict(price=np.array(o.price[:nA]), portfolio=np.array(o.port[:nA + 1]),
                    timestamp=o.timestamp, reward=o.reward, done=bool(o.done),
                    transactionPrice=np.array(o.tp[:nA]), transactionUnits=np.array(o.tu[:nA]),
                    transactionCost=np.array(o.tc[:nA]), riskInfo=np.array(o.risk[:nA]),
                    marginCall=bool(o.margin_call), agent_reward=np.array(o.agent_reward[:nA]),
                    shaped=np.array([list(o.shaped[k][:nA]) for k in range(o.n_popped)]).reshape(o.n_popped, nA),
                    n_popped=o.n_popped)

    def set_ctor_uniforms(self, u):
        """canonical uniforms the next reset() of a SINEDYNAMIC* source uses for (freq, mu, amp) per component"""
        cu = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
        self.L.orc_set_ctor_uniforms(C.byref(self.e), _ptr(cu), 0 if cu is None else cu.size)

    def reset(self, normals=None, uniforms=None):
        o = OrcStepOut()
        n = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        self.L.orc_reset(C.byref(self.e), _ptr(n), _ptr(u), C.byref(o))
        return self._out(o)

    def step(self, units=None, asset_idx=None, normals=None, uniforms=None):
        o = OrcStepOut()
        n = None if normals is None else np.ascontiguousarray(normals, dtype=np.float64)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        if units is None:
            mode, un, ai = A.MODE_HOLD, None, 0
        elif asset_idx is not None:
            mode, un, ai = A.MODE_SINGLE, np.array([units], dtype=np.float64), int(asset_idx)
        else:
            mode, un, ai = A.MODE_MULTI, np.ascontiguousarray(units, dtype=np.float64), 0
        self.L.orc_step(C.byref(self.e), mode, _ptr(un), ai, _ptr(n), _ptr(u), C.byref(o))
        return self._out(o)

    def shaper_feed(self, raw, port, done):
        """Feed one raw reward vector to the n-step shaper; returns (popped rewards (n_popped, ra), n_popped)."""
        raw = np.ascontiguousarray(raw, dtype=np.float64)
        port = np.ascontiguousarray(port, dtype=np.float64)
        o = OrcStepOut()
        self.L.orc_shaper_feed(C.byref(self.e), _ptr(raw), len(raw), _ptr(port), int(done), C.byref(o))
        ra = len(raw)
        return np.array([list(o.shaped[k][:ra]) for k in range(o.n_popped)]).reshape(o.n_popped, ra), o.n_popped

    # -- Portfolio / Broker primitives
    def action_units(self, actions, action_atoms, unit_size):
        """DQN.action_to_transaction (dqn.py:160-179) on the current portfolio: discrete actions -> units"""
        a = np.ascontiguousarray(actions, dtype=np.int8)
        out = np.zeros(self.nA)
        self.L.orc_action_units(C.byref(self.e), a.ctypes.data_as(C.c_void_p), int(action_atoms), float(unit_size),
                                out.ctypes.data_as(C.c_void_p))
        return out

    def weight_units(self, weights):
        """DDPG.action_to_transaction (ddpg.py:182-207) on the current portfolio: fp32 target weights (cash first) -> units"""
        w = np.ascontiguousarray(weights, dtype=np.float32)
        out = np.zeros(self.nA)
        self.L.orc_weight_units(C.byref(self.e), w.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    def handleTransaction(self, i, price, units, cost=0.):
        self.L.orc_handle_transaction(C.byref(self.e), i, price, units, cost)

    def brokerTransaction(self, i, units):
        tp, tu, tc = C.c_double(), C.c_double(), C.c_double()
        risk = self.L.orc_broker_transaction(C.byref(self.e), i, units, C.byref(tp), C.byref(tu), C.byref(tc))
        return dict(transactionPrice=tp.value, transactionUnits=tu.value, transactionCost=tc.value, riskInfo=risk)

    def checkRisk(self, i=None, units=None):
        if i is None:
            return self.L.orc_check_risk(C.byref(self.e))
        return self.L.orc_check_risk_asset(C.byref(self.e), i, units)

    @property
    def prices(self): return np.ctypeslib.as_array(self.e.price)[:self.nA]  # live view, like the reference
    @property
    def ledger(self): return np.ctypeslib.as_array(self.e.ledger)[:self.nA]
    @property
    def meanEntryPrices(self): return np.ctypeslib.as_array(self.e.mep)[:self.nA]
    @property
    def borrowedMarginLedger(self): return np.ctypeslib.as_array(self.e.bm)[:self.nA]
    @property
    def cash(self): return self.e.cash
    @property
    def timestamp(self): return self.e.timestamp
    @property
    def equity(self): return self.L.orc_equity(C.byref(self.e))
    @property
    def assetValue(self): return self.L.orc_asset_value(C.byref(self.e))
    @property
    def pnl(self): return self.L.orc_pnl(C.byref(self.e))
    @property
    def balance(self): return self.L.orc_balance(C.byref(self.e))
    @property
    def availableMargin(self): return self.L.orc_available_margin(C.byref(self.e))
    @property
    def usedMargin(self): return self.L.orc_used_margin(C.byref(self.e))
    @property
    def borrowedMargin(self): return self.L.orc_borrowed_margin(C.byref(self.e))
    @property
    def borrowedAssetValue(self): return self.L.orc_borrowed_asset_value(C.byref(self.e))

    def _vec(self, fn, n):
        out = (C.c_double * (A.MDG_MAX_ASSETS + 1))()
        fn(C.byref(self.e), out)
        return np.array(out[:n])

    @property
    def ledgerNormed(self): return self._vec(self.L.orc_ledger_normed, self.nA)
    @property
    def ledgerNormedFull(self): return self._vec(self.L.orc_ledger_normed_full, self.nA + 1)
    @property
    def ledgerAbsNormed(self): return self._vec(self.L.orc_ledger_abs_normed, self.nA)
    @property
    def ledgerAbsNormedFull(self): return self._vec(self.L.orc_ledger_abs_normed_full, self.nA + 1)


class OracleBatch:
    """N independent oracle envs behind the same tensor layout as the CUDA path."""

    def __init__(self, n_envs, params, reward=None, window=64, seed=0, env_offset=0, threads=1):
        self.L = lib()
        params, self._ext = host_params(params)
        self.N, self.P, self.k = int(n_envs), params, int(window)
        self.R = reward if reward is not None else A.MdgReward(shaper=A.SHAPER_OFF, nstep=1)
        self.nA = params.n_assets
        self.ra = 1 if self.R.reduce_rewards else self.nA
        self.seed, self.env_offset, self.threads = seed, env_offset, threads
        self.h = self.L.orc_batch_create(self.N, C.byref(params), C.byref(self.R), seed, env_offset)
        N, nA, k = self.N, self.nA, self.k
        f8 = np.float64
        self.obs_price = np.zeros((k, nA, N), f8)
        self.obs_port = np.zeros((k, nA + 1, N), f8)
        self.reward = np.zeros(N, f8)
        self.done = np.zeros(N, np.uint8)
        self.trans_price = np.zeros((nA, N), f8)
        self.trans_units = np.zeros((nA, N), f8)
        self.trans_cost = np.zeros((nA, N), f8)
        self.risk = np.zeros((nA, N), np.uint8)
        self.margin_call = np.zeros(N, np.uint8)
        self.agent_reward = np.zeros((self.ra, N), f8)
        self.shaped_reward = np.zeros((self.R.nstep, self.ra, N), f8)
        self.n_popped = np.zeros(N, np.int32)
        self.head = 0
        self.n_valid = 0

    def __del__(self):
        try:
            self.L.orc_batch_destroy(self.h)
        except Exception:
            pass

    def _io(self, units=None, normals=None, uniforms=None):
        io = A.MdgStepIO()
        self._keep = [np.ascontiguousarray(x, dtype=np.float64) if x is not None else None
                      for x in (units, normals, uniforms)]
        io.units, io.normals, io.uniforms = (_ptr(x) for x in self._keep)
        for name in ("obs_price", "obs_port", "reward", "done", "trans_price", "trans_units",
                     "trans_cost", "risk", "margin_call", "agent_reward", "shaped_reward", "n_popped"):
            setattr(io, name, _ptr(getattr(self, name)))
        return io

    def _launch(self, mode=A.MODE_MULTI, asset_idx=0):
        return A.MdgLaunch(n_envs=self.N, env_offset=self.env_offset, seed=self.seed, window=self.k,
                           head=self.head, mode=mode, asset_idx=asset_idx)

    def set_margins(self, required=None, maintenance=None):
        for i in range(self.N):
            e = self.L.orc_batch_env(self.h, i).contents
            if required is not None:
                e.P.required_margin = required
            if maintenance is not None:
                e.P.maintenance_margin = maintenance

    def reset(self, mask=None, fill_ticks=1, clear_nstep=True, normals=None, uniforms=None):
        """normals/uniforms: [fill_ticks][n_slots][N] or None (Philox)."""
        if mask is None:
            self.n_valid = min(self.k, fill_ticks)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        io = self._io(None, normals, uniforms)
        L = self._launch(A.MODE_HOLD)
        self.L.orc_batch_reset(self.h, C.byref(io), C.byref(L), _ptr(m), fill_ticks, int(clear_nstep), self.threads)

    def step(self, units=None, asset_idx=None, normals=None, uniforms=None):
        """units: (N,nA) for the multi-asset step, (N,) with asset_idx, None = hold."""
        self.head = (self.head + 1) % self.k
        self.n_valid = min(self.k, self.n_valid + 1)
        mode = A.MODE_HOLD if units is None else (A.MODE_SINGLE if asset_idx is not None else A.MODE_MULTI)
        io = self._io(units, normals, uniforms)
        L = self._launch(mode, 0 if asset_idx is None else int(asset_idx))
        self.L.orc_batch_step(self.h, C.byref(io), C.byref(L), self.threads)

    def step_actions(self, actions, action_atoms, unit_size, normals=None, uniforms=None):
        """actions: (N,nA) int8 discrete actions -> units as DQN.action_to_transaction (dqn.py:160-179), then step."""
        self.head = (self.head + 1) % self.k
        self.n_valid = min(self.k, self.n_valid + 1)
        io = self._io(None, normals, uniforms)
        acts = np.ascontiguousarray(actions, dtype=np.int8)
        self._keep.append(acts)
        io.actions = acts.ctypes.data_as(C.c_void_p)
        L = self._launch(A.MODE_MULTI)
        L.action_atoms, L.unit_size = int(action_atoms), float(unit_size)
        self.L.orc_batch_step(self.h, C.byref(io), C.byref(L), self.threads)

    def step_weights(self, weights, normals=None, uniforms=None):
        """weights: (N,nA+1) fp32 target weights -> units as DDPG.action_to_transaction (ddpg.py:182-207), then step."""
        self.head = (self.head + 1) % self.k
        self.n_valid = min(self.k, self.n_valid + 1)
        io = self._io(None, normals, uniforms)
        w = np.ascontiguousarray(weights, dtype=np.float32)
        self._keep.append(w)
        io.weights = w.ctypes.data_as(C.c_void_p)
        self.L.orc_batch_step(self.h, C.byref(io), C.byref(self._launch(A.MODE_MULTI)), self.threads)

    def state(self):
        N, nA = self.N, self.nA
        f8 = np.float64
        s = dict(price=np.zeros((nA, N), f8), ledger=np.zeros((nA, N), f8), mean_entry=np.zeros((nA, N), f8),
                 borrowed=np.zeros((nA, N), f8), cash=np.zeros(N, f8),
                 gstate=np.zeros((max(self.P.n_gstate, 1), N), f8), timestamp=np.zeros(N, np.int64),
                 shaper_A=np.zeros((self.ra, N), f8), shaper_B=np.zeros((self.ra, N), f8),
                 nstep_len=np.zeros(N, np.int32))
        st = A.MdgState()
        for k_, v in s.items():
            setattr(st, k_, _ptr(v))
        self.L.orc_batch_export_state(self.h, C.byref(st))
        return s

    def derived(self):
        N, nA = self.N, self.nA
        f8 = np.float64
        d = {n: np.zeros(N, f8) for n in ("equity", "asset_value", "pnl", "balance", "available_margin",
                                          "used_margin", "borrowed_margin", "borrowed_asset_value")}
        d["risk"] = np.zeros(N, np.uint8)
        for n in ("position_values", "pnl_positions", "ledger_normed", "ledger_abs_normed"):
            d[n] = np.zeros((nA, N), f8)
        for n in ("ledger_normed_full", "ledger_abs_normed_full", "position_values_full", "ledger_full"):
            d[n] = np.zeros((nA + 1, N), f8)
        dd = A.MdgDerived()
        for k_, v in d.items():
            setattr(dd, k_, _ptr(v))
        self.L.orc_batch_derived(self.h, C.byref(dd))
        return d

    def time_window(self, n_valid=None):
        """(N, n_valid) timestamps of the window rows (every row is exactly one generator tick)."""
        nv = self.n_valid if n_valid is None else n_valid
        ts = self.state()["timestamp"]
        return ts[:, None] - np.arange(nv - 1, -1, -1)[None, :]

    def port_window(self, n_valid=None):
        nv = self.n_valid if n_valid is None else n_valid
        idx = [(self.head - (nv - 1 - s)) % self.k for s in range(nv)]
        return np.ascontiguousarray(self.obs_port[idx].transpose(2, 0, 1))

    def window(self, n_valid=None):
        """price window (N, n_valid, nA) oldest first, as StackerDiscrete.current_data (un-normalised)."""
        nv = self.n_valid if n_valid is None else n_valid
        idx = [(self.head - (nv - 1 - s)) % self.k for s in range(nv)]
        return np.ascontiguousarray(self.obs_price[idx].transpose(2, 0, 1))


class OracleRewardNorm:
    """N streaming reward normalisers (reward_normalization.pyx) on numpy state laid out like the CUDA path's."""
    KINDS = {"NullShaper": A.RN_NULL, "SharpeFixedWindow": A.RN_SHARPE_FIXED, "SortinoFixedWindowA": A.RN_SORTINO_A,
             "SortinoFixedWindowB": A.RN_SORTINO_B, "SortinoFixedWindowC": A.RN_SORTINO_C,
             "SharpeEWMA": A.RN_SHARPE_EWMA}

    def __init__(self, kind, n_envs, window=2):
        self.L = lib()
        self.N = int(n_envs)
        w = max(int(window), 1)
        self.state = dict(buffer=np.zeros((w, self.N)), size=np.zeros(self.N, np.int32), front=np.zeros(self.N, np.int32),
                          count=np.zeros(self.N, np.int32), mean_est=np.zeros(self.N), ssq=np.zeros(self.N),
                          ewma=np.zeros(self.N), ewma_old=np.zeros(self.N), ewssq_old=np.zeros(self.N),
                          ewssq=np.zeros(self.N), w1=np.ones(self.N), w2=np.ones(self.N))
        self.rn = A.MdgRewardNorm(kind=self.KINDS[kind], window=w, n_envs=self.N, alpha=2 / (float(w + 1)))
        for k_, v in self.state.items():
            setattr(self.rn, k_, _ptr(v))

    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.L.orc_reward_norm_reset(C.byref(self.rn), _ptr(m))

    def stream(self, reward, reset_mask=None):
        r = np.ascontiguousarray(reward, dtype=np.float64)
        m = None if reset_mask is None else np.ascontiguousarray(reset_mask, dtype=np.uint8)
        out = np.zeros(self.N)
        self.L.orc_reward_norm_stream(C.byref(self.rn), _ptr(r), _ptr(m), _ptr(out))
        return out
