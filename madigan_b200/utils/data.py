"""Data containers handed to agents (reference: madigan/utils/data.py:6-61, field names kept).

Every field gains a leading env axis: ``State.price (N,nF)``, ``State.portfolio (N,nA+1)``,
``State.timestamp (N,)``; windows from the preprocessor are ``(N,k,nF)`` / ``(N,k,nA+1)`` / ``(N,k)``.
"""
from dataclasses import dataclass, field
from typing import Any


@dataclass
class State:
    price: Any
    portfolio: Any
    timestamp: Any
    # back-reference used by the preprocessor to find the env-owned observation ring (not part of the
    # reference's State; excluded from comparisons and repr)
    _ring: Any = field(default=None, repr=False, compare=False)


@dataclass
class StateRecurrent:
    price: Any
    portfolio: Any
    timestamp: Any
    action: Any
    reward: Any


@dataclass
class SARSD:
    state: State
    action: Any
    reward: Any
    next_state: State
    done: Any


@dataclass
class SARSDR:
    state: StateRecurrent
    action: Any
    reward: Any
    next_state: StateRecurrent
    done: Any


@dataclass
class BrokerResponse:
    """reference: BrokerResponse<PriceVector>, environments/cpp/DataTypes.h:102-135."""
    event: str
    timestamp: Any
    transactionPrice: Any
    transactionUnits: Any
    transactionCost: Any
    riskInfo: Any
    marginCall: Any


@dataclass
class EnvInfo:
    """reference: EnvInfo<T>, environments/cpp/DataTypes.h:140-149."""
    brokerResponse: BrokerResponse
    dataEnd: bool = False
