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
  if (kind == 3 || (kind >= 4 && kind <= 6) || kind == 9) {
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
