"""SineAdder / SineDynamic / SineDynamicTrend: host-side config handling and the oracle's generator against an
independent numpy restatement (reference: madigan/environments/cpp/DataSource.cpp:581-673, 677-841, 850-1047;
WaveTableOsc.h:113-157; Config.cpp:216-316).  The bit-for-bit comparison with the reference's compiled sources is
tests/test_oracle_vs_reference.py::test_sine_sources_bit_exact_vs_reference; the CUDA path is compared with the
oracle in tests/test_gpu_parity.py."""
import math

import numpy as np
import pytest

from madigan_b200 import _abi as A
from madigan_b200.environments.data_source import ConfigError, make_params
from oracle.oracle import OracleEnv

DYN = dict(freqRange=[[.1, 1., .01], [0.3, 3.0, .01], [5., 15., .1], [10., 50., .1]],
           muRange=[[1., 5., .02], [.3, 3., .05], [.2, 5., .02], [.5, 5., .02]],
           ampRange=[[1., 5., .01], [.3, 3., .02], [.2, 2., .04], [.5, 5., .05]], dX=0.01, noise=.3)
TREND = dict(DYN, trendRange=[[5, 40], [10, 30]], trendIncr=[0.01, 0.02], trendProb=[.02, .05])


def test_generator_table_layout():
    P, names = make_params("SineAdder")
    assert names == ["multi_sine"] and P.n_assets == 1  # ONE asset (DataSource.cpp:655)
    g = P.gen[0]
    assert g.type == A.GEN_SINEADDER and int(g.p[0]) == 4 and P.n_normals == 4 and P.n_gstate == 4
    assert list(P.ext_host[:4]) == [1., 2., 1., 0.]  # freq, mu, amp, phase of component 0

    P, names = make_params("SineDynamic", DYN)
    g = P.gen[0]
    assert names == ["sine_dynamic"] and g.type == A.GEN_SINEDYNAMIC
    assert (P.n_gstate, P.n_normals, P.n_uniforms) == (16, 1, 1)
    assert g.p[2] == 100.  # sampleRate = (int)(1/dX)
    ext = P.ext_host
    # component 0: base frequency .1 -> maxHarms = int(100/(3*.1)+.5) = 333 -> 512 -> tableLen 2048, 9 tables
    rec = ext[:12]
    assert list(rec[:9]) == DYN["freqRange"][0] + DYN["muRange"][0] + DYN["ampRange"][0]
    nt, tl = int(rec[9]), int(rec[10])
    assert nt == 9
    tops = [ext[tl + 3 * t] for t in range(nt)]
    assert tops[0] == .1 * 2. / 100 and all(tops[t + 1] == tops[t] * 2 for t in range(nt - 1))
    ln, off = int(ext[tl + 1]), int(ext[tl + 2])
    assert ln == 2048 and all(int(ext[tl + 3 * t + 1]) == 2048 and int(ext[tl + 3 * t + 2]) == off for t in range(nt))
    tab = ext[off: off + ln + 1]
    assert tab[0] == 0. and tab[ln] == tab[0] and tab[ln // 4] == math.sin(float(ln // 4) * 2. * math.pi / ln)

    P, names = make_params("SineDynamicTrend", TREND)
    g = P.gen[0]
    assert names == ["sine_dynamic_trend"] and (P.n_gstate, P.n_normals, P.n_uniforms) == (16 + 1 + 2, 1, 1 + 2 * 2)
    t_off = int(g.p[5])
    assert int(g.p[4]) == 2 and list(P.ext_host[t_off: t_off + 8]) == [5., 40., .01, .02, 10., 30., .02, .05]


def test_composite_mixes_sine_sources_and_shares_wave_tables():
    cfg = {"a": {"data_source_type": "SineDynamic", "data_source_config": DYN},
           "b": {"data_source_type": "SineDynamicTrend", "data_source_config": TREND},
           "c": {"data_source_type": "OUPair", "data_source_config": {"theta": .015, "phi": .01, "noise": .03}},
           "d": {"data_source_type": "SineAdder"}}
    P, names = make_params("Composite", cfg)
    assert names == ["sine_dynamic", "sine_dynamic_trend", "OUPair_0", "OUPair_1", "multi_sine"]
    one, _ = make_params("SineDynamic", DYN)
    # the second source reuses the first one's four wave tables: only records + lists + trends + SineAdder rows are added
    assert P.n_gen_ext < 2 * one.n_gen_ext
    assert P.gen[1].gslot == 16 and P.gen[2].gslot == 16 + 19 and P.gen[4].gslot == 16 + 19 + 1
    assert P.gen[1].uslot == 1 and P.gen[4].nslot == P.n_normals - 4


def test_config_errors():
    with pytest.raises(ValueError):  # std::length_error, DataSource.cpp:596-599
        make_params("SineAdder", dict(freq=[1., 2.], mu=[1.], amp=[1., 1.], phase=[0., 0.], dX=.01))
    with pytest.raises(ConfigError):  # Config.cpp:217-226
        make_params("SineDynamic", {k: v for k, v in DYN.items() if k != "muRange"})
    with pytest.raises(ConfigError):  # noise is required for the trend source, Config.cpp:262-272
        make_params("SineDynamicTrend", {k: v for k, v in TREND.items() if k != "noise"})
    with pytest.raises(RuntimeError, match="nyquist"):  # std::logic_error, DataSource.cpp:758-765
        make_params("SineDynamic", dict(DYN, dX=0.05))
    with pytest.raises(ValueError):  # :741-743 / :948-953
        make_params("SineDynamic", dict(DYN, ampRange=DYN["ampRange"][:3]))
    with pytest.raises(ValueError):
        make_params("SineDynamicTrend", dict(TREND, trendIncr=[.1]))
    with pytest.raises(ValueError):
        make_params("SineAdder", dict(freq=[1.] * 17, mu=[1.] * 17, amp=[1.] * 17, phase=[0.] * 17, dX=.01))


def _numpy_sine_dynamic(cfg, start, bools, z, trend_u=None):
    """Independent restatement of SineDynamic(Trend)::getData with explicit loops over plain Python floats."""
    K = len(cfg["freqRange"])
    sr = int(1. / cfg["dX"])
    freq, mu, amp = (list(start[j::3]) for j in range(3))
    phasor = [0.] * K
    tables = []
    for c in range(K):
        mh = int(sr / (3.0 * cfg["freqRange"][c][0]) + 0.5)
        ln = (1 << (mh - 1).bit_length()) * 4
        tables.append([math.sin(float(i) * 2. * math.pi / ln) for i in range(ln)] + [0.])
        tables[c][ln] = tables[c][0]
    T = len(cfg.get("trendProb", []))
    tc, trending, direction, length = 1., [False] * T, [1] * T, [0] * T
    out = []
    for t in range(len(z)):
        b = bools[t]
        s = 0.
        for c in range(K):
            def walk(v, up, r):
                return max(r[0], min(r[1], v + (r[2] if up else -r[2])))
            mu[c] = walk(mu[c], b[3 * c], cfg["muRange"][c])
            amp[c] = walk(amp[c], b[3 * c + 1], cfg["ampRange"][c])
            freq[c] = walk(freq[c], b[3 * c + 2], cfg["freqRange"][c])
            phasor[c] += freq[c] / sr
            if phasor[c] >= 1.:
                phasor[c] -= 1.
            ln = len(tables[c]) - 1
            x = phasor[c] * ln
            ip = int(x)
            o = tables[c][ip] + (tables[c][ip + 1] - tables[c][ip]) * (x - ip)
            s += (tc if T else 1.) * (mu[c] + amp[c] * o) if T else mu[c] + amp[c] * o
        if not T:
            out.append(s + z[t] * cfg["noise"])
            continue
        for j in range(T):
            if trending[j]:
                tc += tc * cfg["trendIncr"][j] * direction[j]
                length[j] -= 1
                if length[j] == 0:
                    trending[j] = False
            else:
                trig, ulen = trend_u[t][2 * j], trend_u[t][2 * j + 1]
                if trig < cfg["trendProb"][j]:
                    trending[j] = True
                    direction[j] = -1 if b[3 * K + j] else 1
                    lo, hi = cfg["trendRange"][j]
                    length[j] = int(lo + math.floor(ulen * (hi - lo + 1.)))
            if tc <= .1:
                direction[j] = 1
            tc = max(0.01, tc)
        out.append(s + tc + tc * (z[t] * cfg["noise"]))
    return out


@pytest.mark.parametrize("kind,cfg", [("SineDynamic", DYN), ("SineDynamicTrend", TREND)])
def test_oracle_matches_numpy_restatement(kind, cfg):
    rng = np.random.default_rng(3)
    K, T, n = len(cfg["freqRange"]), len(cfg.get("trendProb", [])), 400
    P, _ = make_params(kind, cfg)
    cu = rng.random(3 * K)
    o = OracleEnv(P, construct=False, ctor_uniforms=cu)
    start = []
    for c in range(K):
        for j, key in enumerate(("freqRange", "muRange", "ampRange")):
            lo, hi = cfg[key][c][:2]
            start.append(lo + cu[3 * c + j] * (hi - lo))
    bools = rng.integers(0, 2, size=(n, 3 * K + T)).astype(bool)
    z = rng.standard_normal(n)
    tu = rng.random((n, 2 * T))
    want = _numpy_sine_dynamic(cfg, start, bools, z, tu)
    for t in range(n):
        bits = sum(1 << (52 - i) for i in range(3 * K + T) if bools[t, i])
        u = np.concatenate([[bits * 2. ** -53], tu[t]])
        got = float(o.tick(normals=[z[t]], uniforms=u)[0])
        assert got == want[t], (t, got, want[t])


def test_sine_adder_closed_form():
    cfg = dict(freq=[1., 0.3, 2.], mu=[2., 2.1, 2.2], amp=[1., 1.2, 1.3], phase=[0., 1., 2.], dX=0.01, noise=0.)
    P, _ = make_params("SineAdder", cfg)
    o = OracleEnv(P, construct=False)
    for t in range(50):
        got = float(o.tick()[0])
        want = sum(m + a * math.sin(2 * math.pi * (ph + t * .01) * f)
                   for f, m, a, ph in zip(cfg["freq"], cfg["mu"], cfg["amp"], cfg["phase"]))
        assert abs(got - want) < 1e-9
