"""On-device episode tearsheet: the batched counterpart of ``madigan/utils/metrics.py`` ``test_summary`` (:83-171).

The reference appends every step of a test episode to Python lists (equity, reward, ledger, transaction cost) and reduces
them with pandas/numba when the episode is over.  Here every env of a slab streams its episode into a few accumulators
in HBM -- one kernel per step -- and ``summary()`` turns them into the same tearsheet fields, per env, without the
episode ever leaving the device.  Field names follow the reference's columns:

    nsteps, mean_equity, final_equity, mean_reward, max_drawdown, mean_transaction_cost, total_transaction_cost,
    equity_returns_offset_<tf>, equity_sharpe_offset_<tf>, equity_sortino_offset_<tf>   (tf = 1, 2, 4, ...),
    time_spent_in_pos_<asset>

(beta coefficients and Q-value means need the agent's outputs / price history and stay on the host side.)
"""
import ctypes as C

import torch

from .. import _abi as A
from .._lib import check, lib


class EpisodeTearsheet:
    def __init__(self, env, n_offsets=6):
        """``n_offsets``: return offsets 1, 2, ..., 2^(n_offsets-1) steps (the reference's integer timeframes,
        metrics.py:71-78)."""
        if not 0 <= int(n_offsets) <= A.MDG_TS_MAX_OFFSETS:
            raise ValueError(f"n_offsets must be in 0..{A.MDG_TS_MAX_OFFSETS}")
        self.env, self._lib = env, lib()
        self.N, self.nA, self.J = env.N, env.nA, int(n_offsets)
        dev = env.device
        z = lambda *s, dt=torch.float64: torch.zeros(s, dtype=dt, device=dev)
        N, nA, J = self.N, self.nA, self.J
        R = 1 << (J - 1) if J else 1
        self.t = dict(nsteps=z(N, dt=torch.int64), active=z(N, dt=torch.uint8), sum_equity=z(N), last_equity=z(N),
                      sum_reward=z(N), peak=z(N), min_valley=z(N), sum_cost=z(N), in_pos=z(nA, N), eq_ring=z(R, N),
                      ret_n=z(max(J, 1), N, dt=torch.int32), ret_mean=z(max(J, 1), N), ret_m2=z(max(J, 1), N),
                      ret_down=z(max(J, 1), N))
        self._ts = A.MdgTearsheet(n_envs=N, n_assets=nA, n_offsets=J, **{k: v.data_ptr() for k, v in self.t.items()})
        self.reset()

    def reset(self, mask=None):
        """Re-arm the envs with ``mask[e]`` set (None: all): empty accumulators, recording from the next update."""
        m = None
        if mask is not None:
            m = torch.as_tensor(mask)
            m = (m.view(torch.uint8) if m.dtype == torch.bool and m.is_cuda else m.to(torch.uint8))
            m = m.to(device=self.env.device, dtype=torch.uint8).contiguous()
        with torch.cuda.device(self.env.device):
            check(self._lib.mdg_tearsheet_reset(C.byref(self._ts), None if m is None else m.data_ptr(), self.env._sptr()))

    def update(self):
        """Record the step the env has just taken.  Call after ``env.step(...)`` and BEFORE the finished envs are reset
        (``env.step(..., auto_reset=True, tearsheet=self)`` does exactly that); an env whose step reported ``done`` is
        recorded one last time and then frozen until ``reset``."""
        env = self.env
        with torch.cuda.device(env.device):
            check(self._lib.mdg_tearsheet_update(C.byref(env.P), C.byref(env._S), C.byref(env._IO), C.byref(self._ts),
                                                 env._sptr()))
        env.launches += 1

    @property
    def active(self):
        return self.t["active"].bool()

    def fields(self):
        names = ["nsteps", "mean_equity", "final_equity", "mean_reward", "max_drawdown", "mean_transaction_cost",
                 "total_transaction_cost"]
        for j in range(self.J):
            tf = 1 << j
            names += [f"equity_returns_offset_{tf}", f"equity_sharpe_offset_{tf}", f"equity_sortino_offset_{tf}"]
        names += [f"time_spent_in_pos_{a}" for a in self.env._asset_names]
        return names

    def summary(self):
        """dict field -> (N,) fp64 tensor, the per-env tearsheet of the episodes recorded so far."""
        names = self.fields()
        out = torch.empty((len(names), self.N), dtype=torch.float64, device=self.env.device)
        with torch.cuda.device(self.env.device):
            check(self._lib.mdg_tearsheet_summary(C.byref(self._ts), out.data_ptr(), self.env._sptr()))
        return {n: out[i] for i, n in enumerate(names)}
