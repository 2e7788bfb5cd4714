"""Batched drop-in for ``madigan/environments/reward_normalization.pyx`` (reference :14-272): streaming reward
normalisers -- one scalar state machine per env, fed one raw reward (a log return) per step.

The reference streams one Python float at a time through a Cython class; here ``stream`` takes the ``(N,)`` reward
tensor of a slab (``env.step(...)[1]``) and runs every env's state machine in one kernel
(``csrc/mdg_rewardnorm.cu``).  Same class names, ``from_config`` keys and factory as the reference:

    SharpeFixedWindow   :61-118   reward / rolling std (Welford add at the head, remove at the tail), min 2 samples
    SortinoFixedWindowA :121-145  the same, negative outputs squared in magnitude
    SortinoFixedWindowB :148-218  std over the below-mean rewards only (negative outputs squared)
    SortinoFixedWindowC :148-218  as B without the squaring
    SharpeEWMA          :221-272  reward / exponentially weighted std, alpha = 2 / (window + 1)
    NullShaper          :49-58    identity
"""
import ctypes as C

import torch

from .. import _abi as A
from .._lib import check, lib


def _cfg_get(cfg, key, default=None):
    if cfg is None:
        return default
    try:
        if key in cfg:
            return cfg[key]
    except TypeError:
        pass
    return getattr(cfg, key, default)


class RewardShaper:
    """``n_envs`` state machines of one kind on ``device``; all tensors are ``[rows][N]`` fp64/int32."""
    KIND = A.RN_NULL

    def __init__(self, window=2, *, n_envs=1, device="cuda", stream=None):
        if not torch.cuda.is_available():
            raise RuntimeError("madigan_b200 reward normalisers need a CUDA device; there is no CPU fallback")
        self._lib = lib()
        self.N = int(n_envs)
        self.window = max(int(window), 1)
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self._stream = stream
        z = lambda *s, dt=torch.float64: torch.zeros(s, dtype=dt, device=self.device)
        N = self.N
        self.t = dict(buffer=z(self.window, N), size=z(N, dt=torch.int32), front=z(N, dt=torch.int32),
                      count=z(N, dt=torch.int32), mean_est=z(N), ssq=z(N), ewma=z(N), ewma_old=z(N), ewssq_old=z(N),
                      ewssq=z(N), w1=torch.ones(N, dtype=torch.float64, device=self.device),
                      w2=torch.ones(N, dtype=torch.float64, device=self.device))
        self._rn = A.MdgRewardNorm(kind=self.KIND, window=self.window, n_envs=N, alpha=2 / (float(self.window + 1)),
                                   **{k_: v.data_ptr() for k_, v in self.t.items()})
        self.reset()

    # reference: RewardShaper.from_config (reward_normalization.pyx:30-35); the batched keys ride along
    @classmethod
    def from_config(cls, config, **kw):
        return cls(_cfg_get(config, "window", 2), **kw)

    def _sptr(self):
        return self._stream.cuda_stream if self._stream is not None else \
            torch.cuda.current_stream(self.device).cuda_stream

    @staticmethod
    def _mask(m, N, device):
        if m is None:
            return None
        m = torch.as_tensor(m)
        if m.dtype == torch.bool:
            m = m.view(torch.uint8) if m.is_cuda else m.to(torch.uint8)
        m = m.to(device=device, dtype=torch.uint8).contiguous()
        if tuple(m.shape) != (N,):
            raise ValueError(f"mask must have shape ({N},)")
        return m

    def reset(self, mask=None):
        """``reset()`` (:84-90 "needs to be called when environment resets / episode ends") for the envs with
        ``mask[e]`` set (None: all)."""
        m = self._mask(mask, self.N, self.device)
        with torch.cuda.device(self.device):
            check(self._lib.mdg_reward_norm_reset(C.byref(self._rn), None if m is None else m.data_ptr(), self._sptr()))

    def stream(self, reward, reset_mask=None, out=None):
        """``out[e] = shaper_e.stream(reward[e])``; envs with ``reset_mask[e]`` set (e.g. the previous step's
        ``done``) are ``reset()`` first, in the same launch."""
        r = reward if (isinstance(reward, torch.Tensor) and reward.dtype == torch.float64 and reward.is_cuda
                       and reward.is_contiguous()) else \
            torch.as_tensor(reward, dtype=torch.float64).to(self.device).contiguous()
        if r.dim() == 0 and self.N == 1:
            r = r.reshape(1)
        if tuple(r.shape) != (self.N,):
            raise ValueError(f"reward must have shape ({self.N},)")
        m = self._mask(reset_mask, self.N, self.device)
        if out is None:
            out = torch.empty(self.N, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            check(self._lib.mdg_reward_norm_stream(C.byref(self._rn), r.data_ptr(), None if m is None else m.data_ptr(),
                                                   out.data_ptr(), self._sptr()))
        return out


class NullShaper(RewardShaper):
    KIND = A.RN_NULL

    @classmethod
    def from_config(cls, config, **kw):
        return cls(1, **kw)


class SharpeFixedWindow(RewardShaper):
    KIND = A.RN_SHARPE_FIXED


class SortinoFixedWindowA(RewardShaper):
    KIND = A.RN_SORTINO_A


class SortinoFixedWindowB(RewardShaper):
    KIND = A.RN_SORTINO_B


class SortinoFixedWindowC(RewardShaper):
    KIND = A.RN_SORTINO_C


class SharpeEWMA(RewardShaper):
    KIND = A.RN_SHARPE_EWMA

    @classmethod
    def from_config(cls, config, **kw):  # the reference reads an ATTRIBUTE here (reward_normalization.pyx:235)
        w = getattr(config, "reward_shape_window", None)
        if w is None:
            w = _cfg_get(config, "reward_shape_window", _cfg_get(config, "window", 2))
        return cls(w, **kw)


def make_reward_normalizer(config, **kw):
    """reference: reward_normalization.pyx:14-22; ``n_envs`` / ``device`` / ``stream`` as keywords."""
    conf = config['reward_shaper_config']
    name = _cfg_get(conf, 'reward_shaper')
    if name in ("None", "none", None):
        return NullShaper(1, **kw)
    cls = globals().get(name)
    if isinstance(cls, type) and issubclass(cls, RewardShaper):
        return cls.from_config(conf, **kw)
    raise NotImplementedError(f"reward_shaper {name} has not been implemented")
