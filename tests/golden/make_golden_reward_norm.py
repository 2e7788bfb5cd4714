"""Golden vectors for the streaming reward normalisers, produced by the REFERENCE's own
Cython module: /root/reference/madigan/environments/reward_normalization.pyx is cythonized + compiled in a temp
directory (Cython and g++ are in the image; the reference tree is read-only and is not copied into the repo) and
its classes are streamed over seeded reward sequences with resets.  Run in the build container only:
    python tests/golden/make_golden_reward_norm.py
Writes tests/golden/reward_norm.npz: rewards (T, L), reset flags (T, L) [reset BEFORE streaming row t], and per case
`<case>` -> outputs (T, L)."""
import os
import shutil
import subprocess
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/madigan/environments/reward_normalization.pyx"
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "NullShaper": ("NullShaper", None),
    "SharpeFixedWindow_w2": ("SharpeFixedWindow", 2),
    "SharpeFixedWindow_w5": ("SharpeFixedWindow", 5),
    "SharpeFixedWindow_w50": ("SharpeFixedWindow", 50),
    "SortinoFixedWindowA_w7": ("SortinoFixedWindowA", 7),
    "SortinoFixedWindowB_w3": ("SortinoFixedWindowB", 3),
    "SortinoFixedWindowB_w20": ("SortinoFixedWindowB", 20),
    "SortinoFixedWindowC_w6": ("SortinoFixedWindowC", 6),
    "SharpeEWMA_w10": ("SharpeEWMA", 10),
    "SharpeEWMA_w50": ("SharpeEWMA", 50),
}


def build_reference(tmp):
    shutil.copy(REF, os.path.join(tmp, "reward_normalization.pyx"))
    with open(os.path.join(tmp, "setup.py"), "w") as f:
        f.write("from setuptools import setup, Extension\nfrom Cython.Build import cythonize\n"
                "setup(ext_modules=cythonize([Extension('reward_normalization', ['reward_normalization.pyx'], "
                "language='c++')], language_level=3))\n")
    subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=tmp, check=True, capture_output=True)
    sys.path.insert(0, tmp)
    import reward_normalization
    return reward_normalization


def make(rn, cls, window):
    if cls == "NullShaper":
        return rn.NullShaper.from_config({})
    if cls == "SharpeEWMA":  # its from_config reads an attribute, not a key (reward_normalization.pyx:235)
        return rn.SharpeEWMA.from_config(types.SimpleNamespace(reward_shape_window=window))
    return getattr(rn, cls).from_config({"window": window})


def main():
    tmp = tempfile.mkdtemp(prefix="mdg_rn_")
    try:
        rn = build_reference(tmp)
        # the factory (reward_normalization.pyx:14-22) resolves names through globals()
        assert type(rn.make_reward_normalizer({"reward_shaper_config": {"reward_shaper": "SharpeFixedWindow", "window": 4}})
                    ).__name__ == "SharpeFixedWindow"
        assert type(rn.make_reward_normalizer({"reward_shaper_config": {"reward_shaper": None}})).__name__ == "NullShaper"
        T, L = 400, 6
        rng = np.random.default_rng(2024)
        rewards = rng.standard_normal((T, L)) * np.array([1., .01, .3, 1e-4, 5., .05])
        rewards[:, 4] -= 1.0  # a lane that is mostly negative
        rewards[rng.random((T, L)) < .03] = 0.  # exact zeros
        resets = rng.random((T, L)) < .01
        resets[0] = False
        out = {}
        for name, (cls, window) in CASES.items():
            res = np.zeros((T, L))
            for lane in range(L):
                sh = make(rn, cls, window)
                for t in range(T):
                    if resets[t, lane]:
                        sh.reset()
                    res[t, lane] = sh.stream(float(rewards[t, lane]))
            out[name] = res
        np.savez(os.path.join(HERE, "reward_norm.npz"), rewards=rewards, resets=resets, **out)
        print("wrote reward_norm.npz:", {k: (float(np.abs(v).max()), int(np.isnan(v).sum())) for k, v in out.items()})
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
