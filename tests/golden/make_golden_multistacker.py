"""Golden vectors for MultiStackerDiscrete, produced by the REFERENCE's own class
(/root/reference/madigan/utils/preprocessor.py:202-288, imported read-only; `np.int`, which that class still uses,
is aliased for numpy 2).  Run in the build container only:
    python tests/golden/make_golden_multistacker.py
Writes tests/golden/multistacker.npz: the streamed price / portfolio rows and, at several stream counts, the
reference's current_data() for every column-wise normaliser."""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
NORMS = [None, "lookback", "lookback_log", "log", "standard_normal", "log_standard_normal"]
K, DILATIONS, NF, T = 6, [1, 2, 4], 3, 40
COUNTS = [24, 25, 26, 27, 31, 40]  # stream counts at which current_data() is recorded (all deques full from 21 on)


def main():
    sys.path.insert(0, REF)
    dummy = types.ModuleType("rollers")
    dummy.Roller = object
    sys.modules.setdefault("rollers", dummy)
    if not hasattr(np, "int"):
        np.int = int
    from madigan.utils import preprocessor
    from madigan.utils.data import State
    rng = np.random.default_rng(5)
    prices = np.abs(10 + np.cumsum(rng.standard_normal((T, NF)) * .3, axis=0)) + .5
    ports = rng.random((T, NF + 1))
    ports /= ports.sum(1, keepdims=True)
    out = dict(prices=prices, ports=ports, k=K, dilations=np.array(DILATIONS), counts=np.array(COUNTS))
    for norm in NORMS:
        st = preprocessor.MultiStackerDiscrete(K, DILATIONS, NF, norm=norm is not None, norm_type=norm or "lookback")
        for t in range(T):
            st.stream_state(State(prices[t], ports[t], np.array(t)))
            if t + 1 in COUNTS:
                cur = st.current_data()
                out[f"price_{norm}_{t + 1}"] = np.array(cur.price)
                out[f"port_{norm}_{t + 1}"] = np.array(cur.portfolio)
                out[f"time_{norm}_{t + 1}"] = np.array(cur.timestamp)
        assert st.feature_output_shape == (K, NF * len(DILATIONS))
    np.savez(os.path.join(HERE, "multistacker.npz"), **out)
    print("wrote multistacker.npz", out["price_None_24"].shape)


if __name__ == "__main__":
    main()
