"""Batched sliding-window preprocessor (reference: madigan/utils/preprocessor.py).

``StackerDiscrete`` keeps the reference's interface (``stream_state``, ``current_data``,
``initialize_history``, ``reset_state``, ``__len__``, ``feature_output_shape``,
``from_config``) but owns no deques: the observation ring lives in HBM next to the env
state and the newest row is written by the step kernel itself.  ``current_data()``
is one launch of the window kernel (ring -> shared-memory tile -> (N,k,nF) with the
normaliser applied), so nothing is copied per step.
"""
import torch

from .. import _abi as A
from .data import State

# make_normalizer's names (reference: utils/preprocessor.py:53-77)
NORM_TYPES = {None: A.NORM_NONE, False: A.NORM_NONE, "none": A.NORM_NONE,
              "lookback": A.NORM_LOOKBACK, "lookback_log": A.NORM_LOOKBACK_LOG, "log": A.NORM_LOG,
              "standard_normal": A.NORM_STANDARD, "log_standard_normal": A.NORM_LOG_STANDARD,
              "expanding": A.NORM_EXPANDING}


def make_normalizer(norm_type):
    if norm_type not in NORM_TYPES or norm_type in (None, False, "none"):
        raise NotImplementedError(f"norm_type {norm_type} is not implemented."
                                  "choose from : 'lookback', 'lookback_log', "
                                  "'standard_normal', 'expanding'")
    return NORM_TYPES[norm_type]


def make_preprocessor(config, n_feats):
    """reference: utils/preprocessor.py:28-50."""
    ptype = config["preprocessor_type"] if isinstance(config, dict) else config.preprocessor_type
    if ptype in ("WindowedStacker", "StackerDiscrete"):
        return StackerDiscrete.from_config(config, n_feats)
    if ptype == "StackerDiscretePairs":
        return StackerDiscretePairs.from_config(config, n_feats)
    if ptype == "StackerDiscreteReturns":
        return StackerDiscreteReturns.from_config(config, n_feats)
    if ptype == "MultiStackerDiscrete":
        return MultiStackerDiscrete.from_config(config, n_feats)
    raise NotImplementedError(f"{ptype} is not implemented ")


class StackerDiscrete:
    """reference: utils/preprocessor.py:143-199."""
    transform = A.XFORM_NONE

    def __init__(self, window_len, n_features, norm=True, norm_type="standard_normal", dtype=torch.float64,
                 channels_first=False):
        self.k = int(window_len)
        self.min_tf = self.k
        self.norm = bool(norm)
        self.norm_type = norm_type if self.norm else None
        if self.norm:
            make_normalizer(norm_type)
        self.dtype = dtype
        self.channels_first = channels_first
        self._feature_output_shape = (self.k, n_features)
        self._env = None
        self._len = 0

    @property
    def feature_output_shape(self):
        return self._feature_output_shape

    @classmethod
    def from_config(cls, config, n_feats):
        get = (lambda c, k_, d=None: c.get(k_, d)) if isinstance(config, dict) else \
              (lambda c, k_, d=None: c[k_] if k_ in c.keys() else d)
        pconf = config["preprocessor_config"] if isinstance(config, dict) else config.preprocessor_config
        norm = get(pconf, "norm", False)
        norm_type = get(pconf, "norm_type", None)
        return cls(get(pconf, "window_length"), n_feats, norm, norm_type)

    def __len__(self):
        return self._len

    def _bind(self, env):
        if self._env is not None and self._env is not env:
            raise ValueError("a StackerDiscrete serves one Env (the ring is env-owned)")
        if env.k != self.k:
            raise ValueError(f"window_length {self.k} != env observation ring length {env.k}")
        self._env = env

    def stream_state(self, state):
        """Append the newest State.  The row is already in the env-owned ring (the step kernel wrote
        it); this only advances the window length, so it costs no memory traffic."""
        ring = getattr(state, "_ring", None)
        if ring is None:
            raise TypeError("StackerDiscrete consumes States produced by madigan_b200 Env.step/reset")
        self._bind(ring[0])
        self._len = min(self.k, self._len + 1)

    def stream(self, data):
        self.stream_state(data[0] if isinstance(data, tuple) else data)

    def current_data(self):
        env = self._env
        if env is None or self._len == 0:
            raise RuntimeError("no data streamed yet")
        price = env.window(self.norm_type, dtype=self.dtype, channels_first=self.channels_first,
                           n_valid=self._len, transform=self.transform)
        port, time = env.portfolio_window(self._len), env.time_window(self._len)
        if self.transform == A.XFORM_RETURNS:  # preprocessor.py:331-332: portfolio[1:], timestamp[1:]
            port, time = port[:, 1:], time[:, 1:]
        return State(price, port, time)

    def current_price(self, out=None, dtype=None, channels_first=None):
        """Price window only (what the agents' nets consume), optionally fp32 / channels-first."""
        env = self._env
        return env.window(self.norm_type, dtype=dtype or self.dtype,
                          channels_first=self.channels_first if channels_first is None else channels_first,
                          n_valid=self._len, out=out, transform=self.transform)

    def initialize_history(self, env):
        """reference: utils/preprocessor.py:191-194 -- no-action steps until the window is full."""
        self._bind(env)
        while len(self) < self.k:
            _state, _reward, _done, _info = env.step()
            self.stream_state(_state)

    def reset_state(self):
        self._len = 0

    def sync(self, env):
        """After ``env.reset(fill_history=True)`` (reset + history fill in one launch)."""
        self._bind(env)
        self._len = env.n_valid


class StackerDiscretePairs(StackerDiscrete):
    """reference: utils/preprocessor.py:295-321 -- the window of the ratio price[:, 0] / price[:, 1] of a
    two-feature source, shape (k, 1), then the normaliser."""
    transform = A.XFORM_PAIR_RATIO

    def __init__(self, window_len, n_features, norm=True, norm_type="standard_normal", **kw):
        assert n_features == 2  # preprocessor.py:301
        super().__init__(window_len, n_features, norm, norm_type, **kw)
        self._feature_output_shape = (self.k, 1)


class StackerDiscreteReturns(StackerDiscrete):
    """reference: utils/preprocessor.py:324-333 -- normaliser, then ``np.diff(price)``, which runs over the LAST
    axis (features): price (k, nF-1); portfolio and timestamp lose their first row."""
    transform = A.XFORM_RETURNS


class MultiStackerDiscrete:
    """reference: utils/preprocessor.py:202-288 -- one window of ``window_len`` rows per dilation d, holding every d-th
    streamed state (the first streamed state, then every d-th after it); ``current_data().price`` is their
    concatenation over the feature axis, (N, k, n_feats * len(dilations)), normalised; portfolio and timestamp come
    from the first dilation.

    No deques here either: the env-owned observation ring already holds every row, so a dilated window is the ring
    read with a stride (one window-kernel launch per dilation, each writing its column block of the output).  The
    env's ring must be deep enough: ``Env(window=...) >= window_len * max(dilations)``.  Column-wise normalisers only
    (``expanding`` runs across the concatenated features in the reference; not supported with several dilations)."""

    def __init__(self, window_len, dilations, n_feats, norm=True, norm_type="standard_normal", dtype=torch.float64):
        self.k = int(window_len)
        self.dilations = [int(d) for d in dilations]
        if not self.dilations or min(self.dilations) < 1:
            raise ValueError("dilations must be positive")
        self.min_tf = self.k
        self.norm = bool(norm)
        self.norm_type = norm_type if self.norm else None
        if self.norm:
            make_normalizer(norm_type)
            if norm_type == "expanding" and len(self.dilations) > 1:
                raise NotImplementedError("'expanding' runs across the concatenated features; not supported here")
        self.max_dilation = max(self.dilations)
        self.n_feats = int(n_feats)
        self.dtype = dtype
        self._feature_output_shape = (self.k, self.n_feats * len(self.dilations))
        self._env = None
        self._count = 0  # states streamed since reset_state

    @property
    def feature_output_shape(self):
        return self._feature_output_shape

    @classmethod
    def from_config(cls, config, n_feats):
        get = (lambda c, k_, d=None: c.get(k_, d)) if isinstance(config, dict) else \
              (lambda c, k_, d=None: c[k_] if k_ in c.keys() else d)
        pconf = config["preprocessor_config"] if isinstance(config, dict) else config.preprocessor_config
        return cls(get(pconf, "window_length"), get(pconf, "dilations"), n_feats, get(pconf, "norm", False),
                   get(pconf, "norm_type", None))

    def _len_of(self, d):
        """rows the reference's deque of dilation d holds after `_count` streamed states (:247-258)"""
        return min(self.k, (self._count + d - 1) // d)

    def __len__(self):
        return self._len_of(self.max_dilation)

    def _bind(self, env):
        if self._env is not None and self._env is not env:
            raise ValueError("a MultiStackerDiscrete serves one Env (the ring is env-owned)")
        if env.k < self.k * self.max_dilation:
            raise ValueError(f"the env's observation ring ({env.k} rows) is shorter than window_length * max dilation "
                             f"= {self.k * self.max_dilation}")
        self._env = env

    def stream_state(self, state):
        ring = getattr(state, "_ring", None)
        if ring is None:
            raise TypeError("MultiStackerDiscrete consumes States produced by madigan_b200 Env.step/reset")
        self._bind(ring[0])
        self._count += 1

    def stream(self, data):
        self.stream_state(data[0] if isinstance(data, tuple) else data)

    def current_data(self):
        env = self._env
        if env is None or self._count == 0:
            raise RuntimeError("no data streamed yet")
        lens = [self._len_of(d) for d in self.dilations]
        if len(set(lens)) != 1:  # np.concatenate of unequal deques raises in the reference too
            raise ValueError("the dilated windows have different lengths; call initialize_history first")
        nv, D, nF = lens[0], len(self.dilations), self.n_feats
        price = torch.empty((env.N, nv, nF * D), dtype=self.dtype, device=env.device)
        for i, d in enumerate(self.dilations):
            age0 = (self._count - 1) % d  # the newest state the deque of dilation d holds
            env.window(self.norm_type, n_valid=nv, out=price, stride=d, age0=age0, out_feat_offset=i * nF)
        d0 = self.dilations[0]
        a0 = (self._count - 1) % d0
        return State(price, env.portfolio_window(nv, stride=d0, age0=a0), env.time_window(nv, stride=d0, age0=a0))

    def initialize_history(self, env):
        """reference :281-284 -- no-action steps until the longest-dilation window is full."""
        self._bind(env)
        while len(self) < self.k:
            _state, _reward, _done, _info = env.step()
            self.stream_state(_state)

    def reset_state(self):
        self._count = 0
