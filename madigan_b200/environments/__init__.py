"""reference: madigan/environments/__init__.py (make_env :9-22, get_env_info :41-56)."""
import copy

from .env import Asset, Env  # noqa: F401
from .data_source import ConfigError, make_params, make_reward  # noqa: F401


def _get(cfg, key, default=None):
    try:
        return cfg[key] if key in cfg else default
    except TypeError:
        return getattr(cfg, key, default)


def make_env(config, test=False, **batched):
    """Same keys as the reference; the batched ones (``n_envs``, ``seed``, ``device``) may be in
    ``config`` or passed as keywords."""
    config = copy.deepcopy(config)
    if test and "data_source_config_test" in config.keys():
        config["data_source_config"] = config["data_source_config_test"]
    if _get(config, "env_type") in ("Synth",):
        env = Env(_get(config, "data_source_type"), _get(config, "init_cash"), config, **batched)
        env.setRequiredMargin(_get(config, "required_margin"))
        env.setMaintenanceMargin(_get(config, "maintenance_margin"))
        env.setTransactionCost(_get(config, "transaction_cost_rel"), _get(config, "transaction_cost_abs"))
        env.setSlippage(_get(config, "slippage_rel"), _get(config, "slippage_abs"))
        return env
    raise NotImplementedError(f"Env type {_get(config, 'env_type')} not implemented")


def get_env_info(env):
    """The 13 accounting fields of the reference, as batched tensors (copies)."""
    return {
        "timestamp": env.timestamp.clone(),
        "riskInfo": env.checkRisk().clone(),
        "prices": env.currentPrices.clone(),
        "equity": env.equity.clone(),
        "cash": env.cash.clone(),
        "pnl": env.pnl.clone(),
        "balance": env.portfolio.balance.clone(),
        "availableMargin": env.availableMargin.clone(),
        "usedMargin": env.usedMargin.clone(),
        "borrowedAssetValue": env.borrowedAssetValue.clone(),
        "borrowedMargin": env.borrowedMargin.clone(),
        "ledger": env.ledger.clone(),
        "ledgerNormed": env.ledgerNormed.clone(),
    }
