// TEST INFRASTRUCTURE, NOT PRODUCT CODE: C entry points around the REFERENCE's own env classes, compiled
// unmodified from /root/reference/madigan/environments/cpp (Env.h, Broker.cpp, Account.cpp, Portfolio.cpp,
// DataSource.cpp, Config.cpp) against oracle/ref_shim.  Used to check oracle/mdg_oracle.c against the real
// reference code and as the `kind: "reference"` CPU baseline.  Built into oracle/_ref/ (git-ignored).
//
// The reference seeds its generators from the wall clock (DataSource.cpp:472,1131,1195,...).  Here the data
// source handed to Env is a subclass that re-seeds the (protected) engine with a known seed, and a clone of
// the engine + distribution objects replays the same draws so that the SAME standard normals can be injected
// into the oracle: libstdc++'s normal_distribution(m, s)(g) is z*s + m with z independent of (m, s).
#include <algorithm>
#include <any>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include <Eigen/Core>
#include <Eigen/Eigen>
#include <highfive/H5Easy.hpp>
#include <highfive/H5File.hpp>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

// TrendOU / TrendyOU keep their engine and state `private`; this test driver has to re-seed the engine and
// read `trending` to know which draws the next getData() will make.  Access specifiers do not change object
// layout with GCC, and the reference's own translation units are compiled without this.
#define private public
#define protected public
#include "Env.h"
#undef private
#undef protected

using namespace madigan;

namespace {

struct OUx : OU {
  using OU::OU;
  void reseed(unsigned s) { generator.seed(s); for (auto& d : noiseDistribution) d.reset(); }
};
struct OUPairx : OUPair {
  using OUPair::OUPair;
  void reseed(unsigned s) { generator.seed(s); ouNoiseDistribution.reset(); randomWalkDistribution.reset(); }
};
struct Synthx : Synth {
  using Synth::Synth;
  void reseed(unsigned s) { generator.seed(s); noiseDistribution.reset(); }
};

// k independent OUPair sources concatenated, as Composite::getData does (DataSource.cpp:439-451).  The
// reference's Composite keys its sub-sources by type name (Config.cpp:117-124), so it cannot hold more than
// one OUPair; the 16-asset benchmark portfolio (8 pairs) is this thin concatenation of reference OUPair objects.
struct MultiPair : DataSourceTick {
  std::vector<std::unique_ptr<OUPairx>> pairs;
  PriceVector data;
  std::size_t ts = 0;
  MultiPair(int k, double theta, double phi, double noise) {
    for (int i = 0; i < k; ++i) {
      pairs.emplace_back(std::make_unique<OUPairx>(theta, phi, noise));
      assets_.push_back(Asset("OUPair_" + std::to_string(2 * i)));
      assets_.push_back(Asset("OUPair_" + std::to_string(2 * i + 1)));
    }
    data.resize(2 * k);
    for (int i = 0; i < 2 * k; ++i) data(i) = 10.;
  }
  const PriceVector& getData() override {
    for (size_t i = 0; i < pairs.size(); ++i) {
      const PriceVector& d = pairs[i]->getData();
      data(2 * i) = d(0);
      data(2 * i + 1) = d(1);
    }
    ts += 1;
    return data;
  }
  const PriceVector& currentData() const override { return data; }
  const PriceVector& currentPrices() const override { return data; }
  int nFeats() const override { return (int)data.size(); }
  void reset() override {
    for (auto& p : pairs) p->reset();
    for (int i = 0; i < (int)data.size(); ++i) data(i) = 10.;
  }
  std::size_t currentTime() const override { return ts; }
};

struct RefEnv {
  std::unique_ptr<Env> env;
  int kind = 0;  // 0 Synth, 1 OU, 2 OUPair, 3 k x OUPair, 4 SimpleTrend, 5 TrendOU, 6 TrendyOU
  SimpleTrend* strend = nullptr;
  TrendOU* trendou = nullptr;
  TrendyOU* trendyou = nullptr;
  std::vector<std::uniform_int_distribution<int>> clen;  // clones of trendLenDist
  int n_assets = 0;
  // clone of the source's RNG objects (same seed, same call order => same draws)
  std::default_random_engine clone;
  std::vector<std::normal_distribution<double>> cn;  // one per distribution object of the source
  OUx* ou = nullptr;
  OUPairx* pair = nullptr;
  Synth* synth = nullptr;
  Gaussian* gauss = nullptr;
  MultiPair* multi = nullptr;
  std::vector<std::default_random_engine> clones;  // kind 3: one engine per pair
  // kinds 10-12: SineAdder, SineDynamic, SineDynamicTrend (one asset, ncomp components)
  SineAdder* sadd = nullptr;
  SineDynamic* sdyn = nullptr;
  SineDynamicTrend* sdt = nullptr;
  int ncomp = 0, ntrend = 0;
};

}  // namespace

extern "C" {

// kind: 0 Synth (p = freq,mu,amp,phase per asset then dX, noise), 1 OU (mean,theta,phi per asset),
//       2 OUPair (theta,phi,noise), 3 n_assets/2 independent OUPairs (theta,phi,noise)
void* ref_env_create(int kind, int n_assets, const double* p, double init_cash, unsigned seed) {
  auto* r = new RefEnv();
  r->kind = kind;
  std::unique_ptr<DataSourceTick> src;
  if (kind == 0 || kind == 7 || kind == 8) {
    std::vector<double> f(n_assets), mu(n_assets), amp(n_assets), ph(n_assets);
    for (int i = 0; i < n_assets; ++i) { f[i] = p[4 * i]; mu[i] = p[4 * i + 1]; amp[i] = p[4 * i + 2]; ph[i] = p[4 * i + 3]; }
    const double dX = p[4 * n_assets], noise = p[4 * n_assets + 1];
    std::unique_ptr<Synth> s;
    if (kind == 0) s = std::make_unique<Synth>(f, mu, amp, ph, dX, noise);
    else if (kind == 7) s = std::make_unique<SawTooth>(f, mu, amp, ph, dX, noise);
    else s = std::make_unique<Triangle>(f, mu, amp, ph, dX, noise);
    r->synth = s.get();
    r->cn.assign(1, std::normal_distribution<double>(0., 1.));
    src = std::move(s);
  } else if (kind == 9) {  // Gaussian: p = mean,var per asset
    std::vector<double> m(n_assets), v(n_assets);
    for (int i = 0; i < n_assets; ++i) { m[i] = p[2 * i]; v[i] = p[2 * i + 1]; }
    auto s = std::make_unique<Gaussian>(m, v);
    s->generator.seed(seed);
    s->timestamp_ = 0;
    r->gauss = s.get();
    r->cn.assign(n_assets, std::normal_distribution<double>(0., 1.));
    src = std::move(s);
  } else if (kind == 10) {  // SineAdder: p = freq,mu,amp,phase per component then dX, noise; n_assets = components
    const int K = n_assets;
    std::vector<double> f(K), mu(K), amp(K), ph(K);
    for (int i = 0; i < K; ++i) { f[i] = p[4 * i]; mu[i] = p[4 * i + 1]; amp[i] = p[4 * i + 2]; ph[i] = p[4 * i + 3]; }
    auto s = std::make_unique<SineAdder>(f, mu, amp, ph, p[4 * K], p[4 * K + 1]);
    s->generator.seed(seed);
    s->noiseDistribution.reset();
    s->timestamp_ = 0;  // left uninitialised by the reference (DataSource.h:284)
    r->sadd = s.get();
    r->ncomp = K;
    n_assets = 1;
    src = std::move(s);
  } else if (kind == 11 || kind == 12) {
    // p = (freq lo,hi,step, mu lo,hi,step, amp lo,hi,step) per component, dX, noise[, T, (lenLo,lenHi,incr,prob) per trend]
    const int K = n_assets;
    std::vector<std::array<double, 3>> fr(K), mr(K), ar(K);
    for (int i = 0; i < K; ++i)
      for (int j = 0; j < 3; ++j) { fr[i][j] = p[9 * i + j]; mr[i][j] = p[9 * i + 3 + j]; ar[i][j] = p[9 * i + 6 + j]; }
    const double dX = p[9 * K], noise = p[9 * K + 1];
    if (kind == 11) {
      auto s = std::make_unique<SineDynamic>(fr, mr, ar, dX, noise);
      s->generator.seed(seed);
      s->noiseDistribution.reset();
      s->timestamp_ = 0;
      s->boolDist.generator.state[0] = 0x9E3779B97F4A7C15ull ^ seed;  // std::random_device in the reference
      s->boolDist.generator.state[1] = 0xD1B54A32D192ED03ull + seed;
      s->boolDist.counter = 0;
      r->sdyn = s.get();
      src = std::move(s);
    } else {
      const int T = (int)p[9 * K + 2];
      std::vector<std::array<int, 2>> tr(T);
      std::vector<double> incr(T), prob(T);
      for (int j = 0; j < T; ++j) {
        const double* q = p + 9 * K + 3 + 4 * j;
        tr[j] = {(int)q[0], (int)q[1]}; incr[j] = q[2]; prob[j] = q[3];
      }
      auto s = std::make_unique<SineDynamicTrend>(fr, mr, ar, tr, incr, prob, dX, noise);
      s->generator.seed(seed);
      s->noiseDistribution.reset();
      s->timestamp_ = 0;
      s->boolDist.generator.state[0] = 0x9E3779B97F4A7C15ull ^ seed;
      s->boolDist.generator.state[1] = 0xD1B54A32D192ED03ull + seed;
      s->boolDist.counter = 0;
      r->sdt = s.get();
      r->ntrend = T;
      src = std::move(s);
    }
    r->ncomp = K;
    n_assets = 1;
  } else if (kind == 1) {
    std::vector<double> m(n_assets), th(n_assets), ph(n_assets);
    for (int i = 0; i < n_assets; ++i) { m[i] = p[3 * i]; th[i] = p[3 * i + 1]; ph[i] = p[3 * i + 2]; }
    auto s = std::make_unique<OUx>(m, th, ph);
    r->ou = s.get();
    r->cn.assign(n_assets, std::normal_distribution<double>(0., 1.));
    src = std::move(s);
  } else if (kind == 2) {
    auto s = std::make_unique<OUPairx>(p[0], p[1], p[2]);
    r->pair = s.get();
    r->cn.assign(2, std::normal_distribution<double>(0., 1.));  // [0] ouNoise, [1] randomWalk
    n_assets = 2;
    src = std::move(s);
  } else if (kind >= 4 && kind <= 6) {
    // per asset: SimpleTrend p = trendProb,minPeriod,maxPeriod,noise,start,dYMin,dYMax;
    //            TrendOU/TrendyOU p = trendProb,minPeriod,maxPeriod,dYMin,dYMax,start,theta,phi,noiseTrend,emaAlpha
    const int np = (kind == 4) ? 7 : 10;
    std::vector<std::vector<double>> c(np, std::vector<double>(n_assets));
    for (int i = 0; i < n_assets; ++i)
      for (int j = 0; j < np; ++j) c[j][i] = p[np * i + j];
    std::vector<int> lo(n_assets), hi(n_assets);
    for (int i = 0; i < n_assets; ++i) { lo[i] = (int)c[1][i]; hi[i] = (int)c[2][i]; }
    if (kind == 4) {
      // definition order (DataSource.cpp:1286-1291) is noise, start, dYMin, dYMax -- the header names them differently (A10)
      auto s = std::make_unique<SimpleTrend>(c[0], lo, hi, c[3], c[4], c[5], c[6]);
      s->generator.seed(seed);
      s->timestamp_ = 0;  // left uninitialised by the reference (DataSource.h:618)
      r->strend = s.get();
      r->cn.assign(n_assets, std::normal_distribution<double>(0., 1.));
      src = std::move(s);
    } else if (kind == 5) {
      auto s = std::make_unique<TrendOU>(c[0], lo, hi, c[3], c[4], c[5], c[6], c[7], c[8], c[9]);
      s->generator.seed(seed);
      r->trendou = s.get();
      r->cn.assign(2 * n_assets, std::normal_distribution<double>(0., 1.));  // [2i] ouNoise, [2i+1] trendNoise
      src = std::move(s);
    } else {
      auto s = std::make_unique<TrendyOU>(c[0], lo, hi, c[3], c[4], c[5], c[6], c[7], c[8], c[9]);
      s->generator.seed(seed);
      r->trendyou = s.get();
      r->cn.assign(n_assets, std::normal_distribution<double>(0., 1.));
      src = std::move(s);
    }
    for (int i = 0; i < n_assets; ++i) r->clen.emplace_back(lo[i], hi[i]);
  } else {
    const int k = n_assets / 2;
    auto s = std::make_unique<MultiPair>(k, p[0], p[1], p[2]);
    r->multi = s.get();
    r->cn.assign(2 * k, std::normal_distribution<double>(0., 1.));  // per pair: [2i] ouNoise, [2i+1] randomWalk
    r->clones.resize(k);
    for (int i = 0; i < k; ++i) {
      s->pairs[i]->reseed(seed + 7919u * (unsigned)i);
      r->clones[i].seed(seed + 7919u * (unsigned)i);
    }
    src = std::move(s);
  }
  r->n_assets = n_assets;
  // Env's own constructor builds a source with the right number of assets (and ticks it once); it is then
  // replaced by the re-seedable one, as Env::setDataSource is meant to be used (Env.h:33-34,167-172)
  if (kind == 3 || (kind >= 4 && kind <= 6) || kind >= 9) {
    std::vector<double> ones((size_t)n_assets, 1.);
    Config inner{{"mean", ones}, {"theta", ones}, {"phi", ones}};
    Config cfg{{"data_source_config", inner}};
    r->env = std::make_unique<Env>("OU", init_cash, cfg);
  } else if (kind == 1) {
    std::vector<double> ones((size_t)n_assets, 1.);
    Config inner{{"mean", ones}, {"theta", ones}, {"phi", ones}};
    Config cfg{{"data_source_config", inner}};
    r->env = std::make_unique<Env>("OU", init_cash, cfg);
  } else if (kind == 0 || kind == 7 || kind == 8) {
    std::vector<double> ones((size_t)n_assets, 1.);
    Config inner{{"freq", ones}, {"mu", ones}, {"amp", ones}, {"phase", ones}, {"dX", 0.01}, {"noise", 0.}};
    Config cfg{{"data_source_config", inner}};
    r->env = std::make_unique<Env>("Synth", init_cash, cfg);
  } else {
    r->env = std::make_unique<Env>("OUPair", init_cash);
  }
  if (r->synth) { r->synth->generator.seed(seed); r->synth->noiseDistribution.reset(); }
  if (r->ou) r->ou->reseed(seed);
  if (r->pair) r->pair->reseed(seed);
  r->clone.seed(seed);
  r->env->setDataSource(std::move(src));
  return r;
}

void ref_env_destroy(void* h) { delete (RefEnv*)h; }

void ref_env_set(void* h, double reqM, double maintM, double tc_rel, double tc_abs, double sl_rel, double sl_abs) {
  auto* r = (RefEnv*)h;
  r->env->setRequiredMargin(reqM);
  r->env->setMaintenanceMargin(maintM);
  r->env->setTransactionCost(tc_rel, tc_abs);
  r->env->setSlippage(sl_rel, sl_abs);
}

// the standard normals the source will consume at its NEXT getData(), in the oracle's slot order
int ref_env_next_normals(void* h, double* z) {
  auto* r = (RefEnv*)h;
  int n = 0;
  if (r->kind == 0 || r->kind == 7 || r->kind == 8) {
    for (int i = 0; i < r->n_assets; ++i) z[n++] = r->cn[0](r->clone);  // one shared distribution, DataSource.cpp:537
  } else if (r->kind == 1 || r->kind == 9) {
    for (int i = 0; i < r->n_assets; ++i) z[n++] = r->cn[i](r->clone);  // one distribution per asset, :1176
  } else if (r->kind == 2) {
    z[n++] = r->cn[1](r->clone);  // random walk of the mean, :1233
    z[n++] = r->cn[0](r->clone);  // x0, :1234
    z[n++] = r->cn[0](r->clone);  // x1, :1235
  } else {
    for (size_t i = 0; i < r->clones.size(); ++i) {
      z[n++] = r->cn[2 * i + 1](r->clones[i]);
      z[n++] = r->cn[2 * i](r->clones[i]);
      z[n++] = r->cn[2 * i](r->clones[i]);
    }
  }
  return n;
}

// trend sources: the draws the next getData() will make, as the oracle's fixed slots -- per asset one standard
// normal z[i] and four uniforms u[4i..4i+3] = (trend-start test, direction, trend length, dY), unused ones 1.0.
// Which draws happen depends on `trending` (and for TrendyOU on the floor test), read from the live object;
// the clone engine makes them in the reference's order (DataSource.cpp:1324-1350, 1457-1493, 1602-1642).
// A trend length L comes back as the u whose oracle mapping a + floor(u (b-a+1)) is L; uniform_real(a,b) is
// a + (b-a) canonical, so its canonical is drawn with param (0,1).  after_reset: the draws of the getData() that
// follows a reset() (which clears `trending`).
int ref_env_next_draws(void* h, double* z, double* u, int after_reset) {
  auto* r = (RefEnv*)h;
  const int n = r->n_assets;
  using NP = std::normal_distribution<double>::param_type;
  std::uniform_real_distribution<double> U(0., 1.);
  auto start_trend = [&](int i, double prob) {
    double rnd = U(r->clone);
    u[4 * i] = rnd;
    if (rnd < prob) {
      u[4 * i + 1] = U(r->clone);
      int L = r->clen[i](r->clone);
      u[4 * i + 2] = ((double)(L - r->clen[i].a()) + 0.5) / (double)(r->clen[i].b() - r->clen[i].a() + 1);
      u[4 * i + 3] = U(r->clone);
    }
  };
  for (int i = 0; i < 4 * n; ++i) u[i] = 1.;
  if (r->kind == 4) {
    SimpleTrend& s = *r->strend;
    for (int i = 0; i < n; ++i) {
      if (after_reset || !s.trending[i]) start_trend(i, s.trendProb[i]);
      z[i] = r->cn[i](r->clone, NP(0., 1.));
    }
  } else if (r->kind == 5) {
    TrendOU& s = *r->trendou;
    for (int i = 0; i < n; ++i) {
      if (!after_reset && s.trending[i]) {
        z[i] = r->cn[2 * i + 1](r->clone, NP(0., 1.));
      } else {
        z[i] = r->cn[2 * i](r->clone, NP(0., 1.));
        start_trend(i, s.trendProb[i]);
      }
    }
  } else if (r->kind == 6) {
    TrendyOU& s = *r->trendyou;
    for (int i = 0; i < n; ++i) {
      z[i] = r->cn[i](r->clone, NP(0., 1.));
      if (!after_reset && s.trending[i]) {
        double tr = s.trendComponent[i];
        tr += tr * (s.dY[i] * s.currentDirection[i]);
        tr = std::max(0.1, tr);
        if (tr <= .1) {
          int L = r->clen[i](r->clone);
          u[4 * i + 2] = ((double)(L - r->clen[i].a()) + 0.5) / (double)(r->clen[i].b() - r->clen[i].a() + 1);
        }
      } else {
        start_trend(i, s.trendProb[i]);
      }
    }
  } else {
    return -1;
  }
  return n;
}

// SineAdder / SineDynamic / SineDynamicTrend: the draws of the next getData() (after a reset() first when
// after_reset), taken from COPIES of the live engine, distributions and boolean generator:
//   z      SineAdder: one standard normal per component; SineDynamic*: one standard normal
//   u      SineDynamic*: u[0] packs the random booleans (boolean b = bit 52-b of u 2^53; 3 per component in the
//          order mu, amp, freq, then bit 3K+j = direction of trend j), u[1+2j] the trend-start test of trend j,
//          u[2+2j] the u whose oracle mapping is the drawn trend length; unused ones 1.0
//   ctor_u the 3K canonical uniforms reset() turns into (freq, mu, amp) per component (after_reset only)
int ref_env_next_sine_draws(void* h, double* z, double* u, double* ctor_u, int after_reset) {
  auto* r = (RefEnv*)h;
  using NP = std::normal_distribution<double>::param_type;
  std::uniform_real_distribution<double> U(0., 1.);
  const int K = r->ncomp;
  if (r->sadd) {
    auto g = r->sadd->generator;
    auto nd = r->sadd->noiseDistribution;
    for (int i = 0; i < K; ++i) z[i] = nd(g, NP(0., 1.));
    return K;
  }
  if (!r->sdyn && !r->sdt) return -1;
  auto g = r->sdyn ? r->sdyn->generator : r->sdt->generator;
  auto nd = r->sdyn ? r->sdyn->noiseDistribution : r->sdt->noiseDistribution;
  auto b = r->sdyn ? r->sdyn->boolDist : r->sdt->boolDist;
  if (after_reset)
    for (int i = 0; i < 3 * K; ++i) ctor_u[i] = U(g);  // freqDist, muDist, ampDist per component, :783-787
  unsigned long long bits = 0;
  for (int i = 0; i < 3 * K; ++i)
    if (b.randBool()) bits |= 1ull << (52 - i);
  const int T = r->ntrend;
  for (int j = 0; j < 1 + 2 * T; ++j) u[j] = 1.;
  if (r->sdt) {
    SineDynamicTrend& s = *r->sdt;
    for (int j = 0; j < T; ++j) {
      if (s.trending[j]) continue;
      const double rnd = U(g);
      u[1 + 2 * j] = rnd;
      if (rnd < s.trendProb[j]) {
        if (b.randBool()) bits |= 1ull << (52 - (3 * K + j));
        auto ld = s.trendLenDist[j];
        const int L = ld(g);
        u[2 + 2 * j] = ((double)(L - ld.a()) + 0.5) / (double)(ld.b() - ld.a() + 1);
      }
    }
  }
  u[0] = (double)bits * 0x1.0p-53;
  z[0] = nd(g, NP(0., 1.));
  return 1;
}

// SineDynamic*: freq, mu, amp, oscillator phase per component [4K], then trendComponent and per trend
// (direction, remaining length, trending) for SineDynamicTrend
void ref_env_sine_state(void* h, double* comp, double* trend_component, int* dir, int* len, int* trending) {
  auto* r = (RefEnv*)h;
  const int K = r->ncomp;
  for (int i = 0; i < K; ++i) {
    if (r->sdyn) { comp[4 * i] = r->sdyn->freq[i]; comp[4 * i + 1] = r->sdyn->mu[i]; comp[4 * i + 2] = r->sdyn->amp[i]; comp[4 * i + 3] = r->sdyn->oscillators[i].phasor; }
    if (r->sdt) { comp[4 * i] = r->sdt->freq[i]; comp[4 * i + 1] = r->sdt->mu[i]; comp[4 * i + 2] = r->sdt->amp[i]; comp[4 * i + 3] = r->sdt->oscillators[i].phasor; }
    if (r->sadd) comp[i] = r->sadd->x[i];
  }
  if (r->sdt) {
    *trend_component = r->sdt->trendComponent;
    for (int j = 0; j < r->ntrend; ++j) { dir[j] = r->sdt->currentDirection[j]; len[j] = r->sdt->currentTrendLen[j]; trending[j] = r->sdt->trending[j]; }
  }
}

// generator internals of the trend sources for state comparison: dY, direction, remaining length, trending
void ref_env_trend_state(void* h, double* dY, int* dir, int* len, int* trending) {
  auto* r = (RefEnv*)h;
  for (int i = 0; i < r->n_assets; ++i) {
    if (r->strend) { dY[i] = r->strend->dY[i]; dir[i] = r->strend->currentDirection[i]; len[i] = r->strend->currentTrendLen[i]; trending[i] = r->strend->trending[i]; }
    if (r->trendou) { dY[i] = r->trendou->dY[i]; dir[i] = r->trendou->currentDirection[i]; len[i] = r->trendou->currentTrendLen[i]; trending[i] = r->trendou->trending[i]; }
    if (r->trendyou) { dY[i] = r->trendyou->dY[i]; dir[i] = r->trendyou->currentDirection[i]; len[i] = r->trendyou->currentTrendLen[i]; trending[i] = r->trendyou->trending[i]; }
  }
}

static void fill_state(RefEnv* r, const State& s, double* price, double* port, long long* ts) {
  for (int i = 0; i < r->n_assets; ++i) price[i] = s.price(i);
  for (int i = 0; i < r->n_assets + 1; ++i) port[i] = s.portfolio(i);
  *ts = (long long)s.timestamp;
}

void ref_env_reset(void* h, double* price, double* port, long long* ts) {
  auto* r = (RefEnv*)h;
  State s = r->env->reset();
  fill_state(r, s, price, port, ts);
}

// Env::step(units) (mode 1), Env::step() (mode 0), Env::step(idx, units) (mode 2)
void ref_env_step(void* h, int mode, const double* units, int asset_idx, double* price, double* port, long long* ts,
                  double* reward, int* done, double* tp, double* tu, double* tc, int* risk, int* margin_call) {
  auto* r = (RefEnv*)h;
  const int n = r->n_assets;
  if (mode == 1) {
    AmountVector u(n);
    for (int i = 0; i < n; ++i) u(i) = units[i];
    auto out = r->env->step(u);
    fill_state(r, std::get<0>(out), price, port, ts);
    *reward = std::get<1>(out);
    *done = std::get<2>(out) ? 1 : 0;
    const auto& br = std::get<3>(out).brokerResponse;
    for (int i = 0; i < n; ++i) {
      tp[i] = br.transactionPrice(i); tu[i] = br.transactionUnits(i); tc[i] = br.transactionCost(i);
      risk[i] = (int)br.riskInfo[i];
    }
    *margin_call = br.marginCall ? 1 : 0;
  } else if (mode == 2) {
    auto out = r->env->step(asset_idx, units[0]);
    fill_state(r, std::get<0>(out), price, port, ts);
    *reward = std::get<1>(out);
    *done = std::get<2>(out) ? 1 : 0;
    const auto& br = std::get<3>(out).brokerResponse;
    for (int i = 0; i < n; ++i) { tp[i] = 0; tu[i] = 0; tc[i] = 0; risk[i] = 0; }
    tp[asset_idx] = br.transactionPrice; tu[asset_idx] = br.transactionUnits; tc[asset_idx] = br.transactionCost;
    risk[asset_idx] = (int)br.riskInfo;
    *margin_call = br.marginCall ? 1 : 0;
  } else {
    auto out = r->env->step();
    fill_state(r, std::get<0>(out), price, port, ts);
    *reward = std::get<1>(out);
    *done = std::get<2>(out) ? 1 : 0;
    for (int i = 0; i < n; ++i) { tp[i] = 0; tu[i] = 0; tc[i] = 0; risk[i] = 0; }
    *margin_call = 0;
  }
}

// accounting properties (Env.h:80-92)
void ref_env_accounting(void* h, double* out /*[8]*/, double* ledger, double* mep) {
  auto* r = (RefEnv*)h;
  Env& e = *r->env;
  out[0] = e.equity(); out[1] = e.cash(); out[2] = e.pnl(); out[3] = e.portfolio()->balance();
  out[4] = e.availableMargin(); out[5] = e.usedMargin(); out[6] = e.borrowedMargin(); out[7] = e.borrowedAssetValue();
  for (int i = 0; i < r->n_assets; ++i) { ledger[i] = e.ledger()(i); mep[i] = e.meanEntryPrices()(i); }
}

// CPU baseline: the reference's single-env step loop, `steps` calls of Env::step(units) cycling over `n_act`
// pre-generated action vectors, with reset on done.  Returns the number of resets.
long long ref_env_run(void* h, long long steps, const double* acts, int n_act) {
  auto* r = (RefEnv*)h;
  const int n = r->n_assets;
  AmountVector u(n);
  long long resets = 0;
  for (long long s = 0; s < steps; ++s) {
    const double* a = acts + (s % n_act) * n;
    for (int i = 0; i < n; ++i) u(i) = a[i];
    auto out = r->env->step(u);
    if (std::get<2>(out)) { r->env->reset(); ++resets; }
  }
  return resets;
}

}  // extern "C"
