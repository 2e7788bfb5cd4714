"""Device-side n-step replay buffer (reference: madigan/utils/buffers/replay_buffer.py + nstep_buffer.py:315-361).

``ReplayBuffer.add`` appends a Python SARSD per env-step and ``_sample`` ``np.stack``s Python objects per sample;
here the transitions of all N envs of a slab are ingested by one kernel per step and sampled by one gather kernel,
so observations never leave HBM between env and agent.  The n-step pops themselves (full-buffer pop, drain on done,
DSR/DDR/... shaping) were already produced by the step kernel (``env.shaped_reward`` / ``env.n_popped``).

Observations are stored once per step (slot ``t % depth``): the price window the agents' nets consume
(``(N, k, nF)``, normalised, fp32 by default) and the newest portfolio row -- the agents read ``portfolio[:, -1]`` only
(dqn.py:188-199).  See include/madigan_b200.h for the record layout and the one documented deviation (the
next_state of a terminal transition)."""
import ctypes as C

import torch

from .. import _abi as A
from .._lib import check, lib
from .data import SARSD, State


class DeviceReplay:
    def __init__(self, env, depth=64, capacity=None, n_action=None, norm_type=None, dtype=torch.float32, seed=0):
        """``depth``: observation slots (steps of history kept); ``capacity``: transitions in the ring (default
        N * (depth - nstep - 1), so that a stored transition's state is normally still resident)."""
        self.env, self._lib = env, lib()
        N, k, nA = env.N, env.k, env.nA
        self.N, self.k, self.nF, self.n_port = N, k, nA, nA + 1
        self.depth = int(depth)
        self.nstep, self.ra = int(env.R.nstep), int(env.ra)
        if env.R.shaper == A.SHAPER_OFF:
            raise ValueError("DeviceReplay needs an Env built with reward=... (the step kernel makes the n-step rewards)")
        if self.depth <= self.nstep + 1:
            raise ValueError("depth must exceed nstep + 1")
        cap_max = N * (self.depth - self.nstep - 1)
        self.capacity = int(capacity) if capacity else cap_max
        if self.capacity > cap_max:  # older records would point at overwritten observation slots
            raise ValueError(f"capacity {self.capacity} exceeds N * (depth - nstep - 1) = {cap_max}")
        self.n_action = int(n_action) if n_action else nA
        self.norm_type, self.dtype, self.seed = norm_type, dtype, int(seed)
        dev = env.device
        z = lambda *s, dt: torch.zeros(s, dtype=dt, device=dev)
        self.obs_price = torch.empty((self.depth, N, k, nA), dtype=dtype, device=dev)
        self.obs_port = torch.empty((self.depth, N, nA + 1), dtype=torch.float64, device=dev)
        self.t = dict(t_env=z(self.capacity, dt=torch.int32), t_state_slot=z(self.capacity, dt=torch.int32),
                      t_next_slot=z(self.capacity, dt=torch.int32),
                      t_state_step=torch.full((self.capacity,), -2 ** 63, dtype=torch.int64, device=dev),  # INT64_MIN = empty
                      t_done=z(self.capacity, dt=torch.uint8), t_reward=z(self.capacity, self.ra, dt=torch.float64),
                      t_action=z(self.capacity, self.n_action, dt=torch.float64),
                      cursor=z(1, dt=torch.int64), act_ring=z(self.nstep, N, self.n_action, dt=torch.float64))
        self._rp = A.MdgReplay(capacity=self.capacity, depth=self.depth, nstep=self.nstep, ra=self.ra,
                               n_action=self.n_action, **{k_: v.data_ptr() for k_, v in self.t.items()})
        self._out = {}
        self.step_count = -1  # index of the last ingested step; observe_start() stores slot -1 % depth
        self.draws = 0

    # ------------------------------------------------------------------ ingest
    def _store_obs(self, step):
        slot = step % self.depth
        self.env.window(self.norm_type, dtype=self.dtype, out=self.obs_price[slot])
        self.obs_port[slot].copy_(self.env.t["obs_port"][self.env.head].t())

    def observe_start(self):
        """Store the observation the first action is taken from (after ``env.reset(fill_history=True)``)."""
        with self.env._copy_ctx():
            self._store_obs(self.step_count)

    def add(self, action):
        """Call after ``env.step(..., auto_reset=True)`` / ``env.step_actions(...)`` with the action that was taken
        ((N, n_action), any real dtype): stores the new observation and appends the transitions that step popped."""
        env = self.env
        self.step_count += 1
        t = self.step_count
        a = action if (isinstance(action, torch.Tensor) and action.dtype == torch.float64 and action.is_cuda
                       and action.is_contiguous()) else \
            torch.as_tensor(action).to(device=env.device, dtype=torch.float64).contiguous()
        if a.shape != (self.N, self.n_action):
            raise ValueError(f"action must have shape {(self.N, self.n_action)}")
        with env._copy_ctx():
            self._store_obs(t)
        T = env.t
        with torch.cuda.device(env.device):
            check(self._lib.mdg_replay_append(
                C.byref(self._rp), self.N, t, a.data_ptr(), T["shaped_reward"].data_ptr(), T["n_popped"].data_ptr(),
                T["nstep_len"].data_ptr() if self.nstep > 1 else None, env._done_u8.data_ptr(), env._sptr()))

    def __len__(self):
        return min(int(self.t["cursor"].item()), self.capacity)

    # ------------------------------------------------------------------ sample
    def sample(self, n):
        """``ReplayBuffer.sample`` (replay_buffer.py:94-103): (SARSD of batched tensors, None).  Runs on the env's
        stream (``Env.bind_stream``).  ``self.last_valid`` flags the rows that hold a real transition (all of them
        unless the ring is nearly empty or mostly stale)."""
        dev = self.env.device
        out = self._out.get(n)
        if out is None:  # output buffers are allocated once per batch size and reused (a sample is consumed by the
            # agent's training step before the next one is drawn; clone() what must outlive that)
            out = self._out[n] = dict(
                idx=torch.empty(n, dtype=torch.int64, device=dev),
                state_price=torch.empty((n, self.k, self.nF), dtype=self.dtype, device=dev),
                next_price=torch.empty((n, self.k, self.nF), dtype=self.dtype, device=dev),
                state_port=torch.empty((n, self.n_port), dtype=torch.float64, device=dev),
                next_port=torch.empty((n, self.n_port), dtype=torch.float64, device=dev),
                action=torch.empty((n, self.n_action), dtype=torch.float64, device=dev),
                reward=torch.empty((n, self.ra), dtype=torch.float64, device=dev),
                done=torch.empty(n, dtype=torch.uint8, device=dev))
        b = A.MdgReplayBatch(**{k_: v.data_ptr() for k_, v in out.items()})
        self.draws += 1
        with torch.cuda.device(dev):
            check(self._lib.mdg_replay_sample(
                C.byref(self._rp), self.N, self.step_count, self.obs_price.data_ptr(), self.obs_port.data_ptr(),
                self.k * self.nF, self.n_port, A.DTYPE_F32 if self.dtype == torch.float32 else A.DTYPE_F64, n,
                self.seed, self.draws, C.byref(b), self.env._sptr()))
        self.last_idx = out["idx"]
        # (batch,) bool: False where 64 draws found no resident transition (the row is zero-filled with done = True)
        with self.env._copy_ctx():
            self.last_valid = out["idx"] >= 0
        sarsd = SARSD(State(out["state_price"], out["state_port"], None), out["action"], out["reward"],
                      State(out["next_price"], out["next_port"], None), out["done"].bool())
        return sarsd, None
