/*
 * mdg_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see mdg_oracle.h).
 *
 * CPU restatement of the reference's Env.step hot path in plain C.  All file:line
 * citations are relative to /root/reference/madigan/.
 *
 * Conventions the reference leaves undefined, fixed here:
 *   - strict IEEE fp64, no FMA contraction (build with -ffp-contract=off); the
 *     reference is built with -ffast-math -mfma (environments/cpp/CMakeLists.txt:5);
 *   - every Eigen dot()/sum() is a left-to-right fold starting from the first term;
 *   - timestamp starts at 0 and is incremented by every getData();
 *   - libstdc++ <random> is replaced by an injected stream with FIXED slots per
 *     asset: normal_distribution(mu,sigma) == z*sigma+mu, uniform_real(a,b) ==
 *     a+u*(b-a), uniform_int(a,b) == a+floor(u*(b-a+1)); when no stream is given
 *     the draws come from Philox4x32-10 keyed by (seed, env id, tick, slot).
 *
 * PARITY STATUS: pinned.
 *   - Ledger (Portfolio/Broker): the reference's own known-answer tests
 *     (environments/cpp/tests/envTest.py:102-566, envTest.cpp:171-266), replayed in
 *     tests/test_oracle_ledger_kat.py.
 *   - Env::step end to end -- all nine generators (Synth, SawTooth, Triangle,
 *     Gaussian, OU, OUPair, SimpleTrend, TrendOU, TrendyOU), reward clamp, done,
 *     slippage/cost, risk gates under leverage, accounting properties: BIT-EXACT
 *     against the reference's own C++ sources compiled here unmodified
 *     (oracle/ref_build.py -> oracle/_ref, tests/test_oracle_vs_reference.py), with
 *     the reference's <random> draws replayed into the injected stream.
 *   - Shapers: the reference's own Python (tests/golden/shapers.npz).
 *   - Not pinned: Composite's asset ORDER (the reference iterates an
 *     unordered_map, DataSource.cpp:418-433; here: config insertion order).
 */
#include "mdg_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define PI2 (3.141592653589793238463 * 2) /* environments/cpp/DataSource.h:24 */

/* ------------------------------------------------------------------ */
/* Philox4x32-10 and draw conventions                                   */
/* ------------------------------------------------------------------ */
void orc_philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* counter = (env id, stream<<16 | block, tick lo, tick hi); each block yields two draws */
static void philox_block(uint64_t seed, int64_t gid, int64_t tick, int stream, int block,
                         uint64_t *x0, uint64_t *x1) {
  uint32_t ctr[4] = {(uint32_t)gid, ((uint32_t)stream << 16) | (uint32_t)block,
                     (uint32_t)(uint64_t)tick, (uint32_t)((uint64_t)tick >> 32)};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  orc_philox4x32_10(ctr, key, o);
  *x0 = ((uint64_t)o[1] << 32) | o[0];
  *x1 = ((uint64_t)o[3] << 32) | o[2];
}

double orc_draw_normal(uint64_t seed, int64_t gid, int64_t tick, int slot) {
  uint64_t x0, x1;
  philox_block(seed, gid, tick, 0, slot >> 1, &x0, &x1);
  double u1 = ((double)(x0 >> 12) + 0.5) * 0x1.0p-52; /* (0,1) */
  double u2 = (double)(x1 >> 11) * 0x1.0p-53;         /* [0,1) */
  double r = sqrt(-2.0 * log(u1));
  double a = PI2 * u2;
  return (slot & 1) ? r * sin(a) : r * cos(a);
}

double orc_draw_uniform(uint64_t seed, int64_t gid, int64_t tick, int slot) {
  uint64_t x0, x1;
  philox_block(seed, gid, tick, 1, slot >> 1, &x0, &x1);
  return (double)(((slot & 1) ? x1 : x0) >> 11) * 0x1.0p-53;
}

/* uniform draws of the SINEDYNAMIC* constructor (stream 3) and reset() (stream 2) */
static double ctor_uniform(const OrcEnv *e, int stream, int asset, int slot) {
  if (e->n_ctor_u > 0) return e->ctor_u[slot];
  uint64_t x0, x1;
  int s = 64 * asset + slot;
  philox_block(e->seed, e->gid, e->timestamp, stream, s >> 1, &x0, &x1);
  return (double)(((s & 1) ? x1 : x0) >> 11) * 0x1.0p-53;
}

typedef struct Draws {
  const OrcEnv *e;
  const double *normals;
  const double *uniforms;
  int64_t tick;
} Draws;

static double dn(const Draws *d, int slot) {
  return d->normals ? d->normals[slot] : orc_draw_normal(d->e->seed, d->e->gid, d->tick, slot);
}
static double du(const Draws *d, int slot) {
  return d->uniforms ? d->uniforms[slot] : orc_draw_uniform(d->e->seed, d->e->gid, d->tick, slot);
}

/* ------------------------------------------------------------------ */
/* Generators  (environments/cpp/DataSource.cpp)                        */
/* ------------------------------------------------------------------ */
static inline int64_t dbl_bits(double x) { int64_t b; memcpy(&b, &x, 8); return b; }
static inline double bits_dbl(int64_t b) { double x; memcpy(&x, &b, 8); return x; }
#define FLAG_TRENDING 1
#define FLAG_DIRPOS 2
static inline int64_t pack_flags(int trending, int dir, int32_t len) {
  return (int64_t)(trending ? FLAG_TRENDING : 0) | (int64_t)(dir > 0 ? FLAG_DIRPOS : 0) |
         (int64_t)((uint64_t)(uint32_t)len << 32);
}

static inline double u_real(double u, double a, double b);
/* freq, mu, amp of every component ~ uniform_real(lo, hi): DataSource.cpp:771-773 / 783-787, :972-974 / 992-996 */
static void sine_dynamic_sample(OrcEnv *e, int i, int stream) {
  const MdgAssetGen *g = &e->P.gen[i];
  double *gs = &e->gstate[g->gslot];
  const int K = (int)g->p[0];
  const double *cp = e->P.gen_ext + (int64_t)g->p[1];
  for (int c = 0; c < K; ++c, cp += 12)
    for (int j = 0; j < 3; ++j)
      gs[4 * c + j] = u_real(ctor_uniform(e, stream, i, 3 * c + j), cp[3 * j], cp[3 * j + 1]);
}

static void gen_start(OrcEnv *e) {
  /* initial generator state = constructor state (initParams of each source) */
  const MdgParams *P = &e->P;
  for (int i = 0; i < P->n_assets; ++i) {
    const MdgAssetGen *g = &P->gen[i];
    double *gs = g->gslot >= 0 ? &e->gstate[g->gslot] : 0;
    switch (g->type) {
      case MDG_GEN_SYNTH: case MDG_GEN_SAWTOOTH: case MDG_GEN_TRIANGLE:
        gs[0] = g->p[3];  /* x = phase, DataSource.cpp:463 */
        e->price[i] = 0.; /* currentData_ resized only, :471 */
        break;
      case MDG_GEN_OU: e->price[i] = g->p[0]; break;          /* :1129 */
      case MDG_GEN_GAUSSIAN: e->price[i] = g->p[0]; break;    /* :1066 */
      case MDG_GEN_OUPAIR:
        e->price[i] = 10.;                                    /* :1192 */
        if (g->role == 0) gs[0] = 10.;                        /* :1194 */
        break;
      case MDG_GEN_SIMPLETREND:
        e->price[i] = g->p[4];                                /* :1273 */
        gs[0] = 0.;                                           /* dY :1276 */
        gs[1] = bits_dbl(pack_flags(0, 1, 0));                /* :1270,1274-1275 */
        break;
      case MDG_GEN_TRENDOU:
        e->price[i] = g->p[5]; /* :1397 (quirk A9 normalised: every asset starts at start[i]) */
        gs[0] = g->p[5];       /* ouMean :1384 */
        gs[1] = 0.;
        gs[2] = bits_dbl(pack_flags(0, 1, 0));
        break;
      case MDG_GEN_TRENDYOU:
        e->price[i] = g->p[5]; /* :1542 */
        gs[0] = 0.;            /* ouComponent :1537 */
        gs[1] = g->p[5];       /* trendComponent :1538 */
        gs[2] = 0.;
        gs[3] = bits_dbl(pack_flags(0, 1, 0));
        break;
      case MDG_GEN_SINEADDER: { /* x = phase, DataSource.cpp:652; currentData_ resized only */
        const int K = (int)g->p[0];
        const double *cp = P->gen_ext + (int64_t)g->p[1];
        for (int c = 0; c < K; ++c) gs[c] = cp[4 * c + 3];
        e->price[i] = 0.;
        break;
      }
      case MDG_GEN_SINEDYNAMIC: case MDG_GEN_SINEDYNAMICTREND: { /* :741-780, :938-989 */
        const int K = (int)g->p[0];
        sine_dynamic_sample(e, i, 3);
        for (int c = 0; c < K; ++c) gs[4 * c + 3] = 0.; /* phasor, WaveTableOsc.h:62 */
        if (g->type == MDG_GEN_SINEDYNAMICTREND) {
          const int T = (int)g->p[4];
          gs[4 * K] = 1.; /* trendComponent :982 */
          for (int j = 0; j < T; ++j) gs[4 * K + 1 + j] = bits_dbl(pack_flags(0, 1, 0)); /* :983-988 */
        }
        e->price[i] = 0.;
        break;
      }
    }
  }
}

static void gen_reset(OrcEnv *e) {
  /* DataSource::reset(): Synth/OU/Gaussian do nothing (DataSource.h:232,466) */
  const MdgParams *P = &e->P;
  for (int i = 0; i < P->n_assets; ++i) {
    const MdgAssetGen *g = &P->gen[i];
    double *gs = g->gslot >= 0 ? &e->gstate[g->gslot] : 0;
    switch (g->type) {
      case MDG_GEN_OUPAIR: /* DataSource.cpp:1242-1246 */
        e->price[i] = 10.;
        if (g->role == 0) gs[0] = 10.;
        break;
      case MDG_GEN_SIMPLETREND: /* :1352-1359 (dY is left as is) */
        e->price[i] = g->p[4];
        gs[1] = bits_dbl(pack_flags(0, 1, 0));
        break;
      case MDG_GEN_TRENDOU: { /* :1495-1502 (direction and dY are left as is) */
        int64_t f = dbl_bits(gs[2]);
        e->price[i] = g->p[5];
        gs[0] = g->p[5];
        gs[2] = bits_dbl(pack_flags(0, (f & FLAG_DIRPOS) ? 1 : -1, 0));
        break;
      }
      case MDG_GEN_TRENDYOU: { /* :1644-1657 */
        int64_t f = dbl_bits(gs[3]);
        gs[0] = 0.;
        gs[1] = g->p[5];
        e->price[i] = g->p[5];
        gs[3] = bits_dbl(pack_flags(0, (f & FLAG_DIRPOS) ? 1 : -1, 0));
        break;
      }
      case MDG_GEN_SINEDYNAMIC: case MDG_GEN_SINEDYNAMICTREND:
        /* :783-787, :992-996: new freq, mu, amp; the oscillators' phase and the trend state are kept */
        sine_dynamic_sample(e, i, 2);
        break;
      default: break;
    }
  }
}

static inline double dmax(double a, double b) { return (a < b) ? b : a; } /* std::max */

/* uniform_int_distribution(a,b) and uniform_real_distribution(a,b) from one u in [0,1) */
static inline int32_t u_int(double u, double a, double b) { return (int32_t)(a + floor(u * (b - a + 1.))); }
static inline double u_real(double u, double a, double b) { return a + u * (b - a); }
static inline double dmin(double a, double b) { return (b < a) ? b : a; } /* std::min */
/* random boolean b of this tick (randomBoolGenerator.h:8-14): bit 52-b of floor(u 2^53) */
static inline int u_bool(uint64_t bits, int b) { return (int)((bits >> (52 - b)) & 1u); }
/* bounded +-step walk of updateParams (DataSource.cpp:804-812) */
static inline double sine_walk(double v, int up, const double *r) {
  return dmax(r[0], dmin(r[1], v + (up ? r[2] : -r[2])));
}
/* WaveTableOsc::setFreq + process (WaveTableOsc.h:77-110) of the component with record cp */
static double osc_process(const double *ext, const double *cp, double incr, double *phasor) {
  const int nt = (int)cp[9];
  const double *tl = ext + (int64_t)cp[10];
  int idx = 0;
  while (incr >= tl[3 * idx] && idx < nt - 1) ++idx; /* setFreq :80-85 */
  *phasor += incr;                                   /* updatePhase :39 */
  if (*phasor >= 1.) *phasor -= 1.;
  const int len = (int)tl[3 * idx + 1];              /* getOutput :89-101 */
  const double *tab = ext + (int64_t)tl[3 * idx + 2];
  const double temp = *phasor * len;
  const int ip = (int)temp;
  const double frac = temp - ip;
  const double s0 = tab[ip], s1 = tab[ip + 1];
  return s0 + (s1 - s0) * frac;
}

void orc_tick(OrcEnv *e, const double *normals, const double *uniforms) {
  const MdgParams *P = &e->P;
  Draws d = {e, normals, uniforms, e->timestamp};
  for (int i = 0; i < P->n_assets; ++i) {
    const MdgAssetGen *g = &P->gen[i];
    const double *p = g->p;
    double *gs = g->gslot >= 0 ? &e->gstate[g->gslot] : 0;
    switch (g->type) {
      case MDG_GEN_SYNTH: { /* DataSource.cpp:535-543 */
        double x = gs[0];
        e->price[i] = dn(&d, g->nslot) * p[5] + p[1] + p[2] * sin(PI2 * x * p[0]);
        gs[0] = x + p[4];
        break;
      }
      case MDG_GEN_SAWTOOTH: { /* :558-567 */
        double x = gs[0], ip;
        e->price[i] = dn(&d, g->nslot) * p[5] + p[1] + p[2] * modf(x * p[0], &ip);
        gs[0] = x + p[4];
        break;
      }
      case MDG_GEN_TRIANGLE: { /* :569-578 */
        double x = gs[0];
        e->price[i] = dn(&d, g->nslot) * p[5] + p[1] + 4 * p[2] / PI2 * asin(sin(PI2 * x / p[0]));
        gs[0] = x + p[4];
        break;
      }
      case MDG_GEN_GAUSSIAN: /* :1108-1114, normal_distribution(mean, var) */
        e->price[i] = dn(&d, g->nslot) * p[1] + p[0];
        break;
      case MDG_GEN_OU: { /* :1173-1180 */
        double x = e->price[i];
        x += (p[1] * (p[0] - x)) + p[0] * p[2] * dn(&d, g->nslot);
        e->price[i] = x;
        break;
      }
      case MDG_GEN_OUPAIR: { /* :1232-1240, draw order rw, x0, x1 */
        double *mean = (g->role == 0) ? gs : &e->gstate[P->gen[g->partner].gslot];
        if (g->role == 0) *mean += *mean * (dn(&d, g->nslot_aux) * p[2]);
        double m = *mean;
        e->price[i] += (p[0] * (m - e->price[i])) + m * (dn(&d, g->nslot) * p[1]);
        break;
      }
      case MDG_GEN_SIMPLETREND: { /* :1324-1350 */
        double y = e->price[i];
        int64_t f = dbl_bits(gs[1]);
        int trending = (int)(f & FLAG_TRENDING), dir = (f & FLAG_DIRPOS) ? 1 : -1;
        int32_t len = (int32_t)((uint64_t)f >> 32);
        if (trending) {
          y += y * gs[0] * dir;
          if (--len == 0) trending = 0;
        } else {
          double r = du(&d, g->uslot);
          if (r < p[0]) {
            trending = 1;
            dir = (du(&d, g->uslot + 1) < 0.5) ? -1 : 1;
            len = u_int(du(&d, g->uslot + 2), p[1], p[2]);
            gs[0] = u_real(du(&d, g->uslot + 3), p[5], p[6]);
          }
        }
        if (y <= .1) dir = 1;
        y += y * (dn(&d, g->nslot) * p[3]);
        y = dmax(0.01, y);
        e->price[i] = y;
        gs[1] = bits_dbl(pack_flags(trending, dir, len));
        break;
      }
      case MDG_GEN_TRENDOU: { /* :1457-1493 */
        double y = e->price[i];
        int64_t f = dbl_bits(gs[2]);
        int trending = (int)(f & FLAG_TRENDING), dir = (f & FLAG_DIRPOS) ? 1 : -1;
        int32_t len = (int32_t)((uint64_t)f >> 32);
        if (trending) {
          y += y * (gs[1] * dir + dn(&d, g->nslot) * p[8]);
          len -= 1;
          if (len == 0) { trending = 0; gs[0] = y; }
          y = dmax(0.01, y);
          if (y <= .1) dir = 1;
        } else {
          double ou_noise = y * (dn(&d, g->nslot) * p[7]);
          double rev = p[6] * (gs[0] - y);
          y += rev + ou_noise;
          double r = du(&d, g->uslot);
          if (r < p[0]) {
            trending = 1;
            dir = (du(&d, g->uslot + 1) < 0.5) ? -1 : 1;
            len = u_int(du(&d, g->uslot + 2), p[1], p[2]);
            gs[1] = u_real(du(&d, g->uslot + 3), p[3], p[4]);
          }
        }
        e->price[i] = y;
        gs[2] = bits_dbl(pack_flags(trending, dir, len));
        break;
      }
      case MDG_GEN_TRENDYOU: { /* :1602-1642 */
        int64_t f = dbl_bits(gs[3]);
        int trending = (int)(f & FLAG_TRENDING), dir = (f & FLAG_DIRPOS) ? 1 : -1;
        int32_t len = (int32_t)((uint64_t)f >> 32);
        double ou = gs[0], tr = gs[1];
        double ou_noise = tr * (dn(&d, g->nslot) * p[7]);
        double rev = p[6] * (-ou);
        ou += rev + ou_noise;
        if (trending) {
          tr += tr * (gs[2] * dir);
          tr = dmax(0.1, tr);
          if (tr <= .1) {
            dir = 1;
            trending = 1;
            len = u_int(du(&d, g->uslot + 2), p[1], p[2]);
          }
          if (--len == 0) trending = 0;
        } else {
          double r = du(&d, g->uslot);
          if (r < p[0]) {
            trending = 1;
            dir = (du(&d, g->uslot + 1) < 0.5) ? -1 : 1;
            len = u_int(du(&d, g->uslot + 2), p[1], p[2]);
            gs[2] = u_real(du(&d, g->uslot + 3), p[3], p[4]);
          }
        }
        gs[0] = ou; gs[1] = tr;
        e->price[i] = ou + tr;
        gs[3] = bits_dbl(pack_flags(trending, dir, len));
        break;
      }
      case MDG_GEN_SINEADDER: { /* :663-673 */
        const int K = (int)p[0];
        const double *cp = P->gen_ext + (int64_t)p[1];
        double sum = 0.;
        for (int c = 0; c < K; ++c, cp += 4) {
          double x = gs[c];
          sum = sum + ((dn(&d, g->nslot + c) * p[3] + cp[1]) + cp[2] * sin(PI2 * x * cp[0]));
          gs[c] = x + p[2];
        }
        e->price[i] = sum;
        break;
      }
      case MDG_GEN_SINEDYNAMIC: case MDG_GEN_SINEDYNAMICTREND: { /* :802-841, :1002-1047 */
        const int trend = g->type == MDG_GEN_SINEDYNAMICTREND;
        const int K = (int)p[0];
        const double *ext = P->gen_ext, *cp = ext + (int64_t)p[1];
        const uint64_t bits = (uint64_t)(du(&d, g->uslot) * 0x1.0p53);
        double tcmp = trend ? gs[4 * K] : 1.;
        double sum = 0.;
        for (int c = 0; c < K; ++c, cp += 12) { /* updateParams (:802-815) then the component sum (:833-838) */
          double *s = gs + 4 * c;
          s[1] = sine_walk(s[1], u_bool(bits, 3 * c), cp + 3);
          s[2] = sine_walk(s[2], u_bool(bits, 3 * c + 1), cp + 6);
          s[0] = sine_walk(s[0], u_bool(bits, 3 * c + 2), cp);
          double o = osc_process(ext, cp, s[0] / p[2], &s[3]);
          sum = trend ? sum + tcmp * (s[1] + s[2] * o) : sum + (s[1] + s[2] * o);
        }
        double z = dn(&d, g->nslot) * p[3] + 0.; /* normal_distribution(0, noise) */
        if (!trend) { e->price[i] = sum + z; break; }
        const int T = (int)p[4];
        const double *tp = ext + (int64_t)p[5];
        for (int j = 0; j < T; ++j, tp += 4) { /* :1021-1040 */
          int64_t f = dbl_bits(gs[4 * K + 1 + j]);
          int trending = (int)(f & FLAG_TRENDING), dir = (f & FLAG_DIRPOS) ? 1 : -1;
          int32_t len = (int32_t)((uint64_t)f >> 32);
          if (trending) {
            tcmp += tcmp * tp[2] * dir;
            if (--len == 0) trending = 0;
          } else {
            double r = du(&d, g->uslot + 1 + 2 * j);
            if (r < tp[3]) {
              trending = 1;
              dir = u_bool(bits, 3 * K + j) ? -1 : 1;
              len = u_int(du(&d, g->uslot + 2 + 2 * j), tp[0], tp[1]);
            }
          }
          if (tcmp <= .1) dir = 1;
          tcmp = dmax(0.01, tcmp);
          gs[4 * K + 1 + j] = bits_dbl(pack_flags(trending, dir, len));
        }
        gs[4 * K] = tcmp;
        e->price[i] = sum + tcmp + tcmp * z; /* :1041 */
        break;
      }
    }
  }
  e->timestamp += 1; /* every getData(): timestamp_ += 1 */
}

/* ------------------------------------------------------------------ */
/* Portfolio  (environments/cpp/Portfolio.cpp)                          */
/* ------------------------------------------------------------------ */
double orc_asset_value(const OrcEnv *e) { /* :180-182 ledger_.dot(currentPrices_) */
  double s = e->ledger[0] * e->price[0];
  for (int j = 1; j < e->P.n_assets; ++j) s = s + e->ledger[j] * e->price[j];
  return s;
}
static double mep_dot_ledger(const OrcEnv *e) { /* :185 meanEntryPrices_.dot(ledger_) */
  double s = e->mep[0] * e->ledger[0];
  for (int j = 1; j < e->P.n_assets; ++j) s = s + e->mep[j] * e->ledger[j];
  return s;
}
double orc_borrowed_margin(const OrcEnv *e) { /* :207-209 borrowedMargin_.sum() */
  double s = e->bm[0];
  for (int j = 1; j < e->P.n_assets; ++j) s = s + e->bm[j];
  return s;
}
double orc_pnl(const OrcEnv *e) { return orc_asset_value(e) - mep_dot_ledger(e); } /* :184-186 */
double orc_balance(const OrcEnv *e) { /* :192-197 */
  double s = e->ledger[0] * (e->mep[0] * (e->ledger[0] < 0. ? 1. : 0.));
  for (int j = 1; j < e->P.n_assets; ++j)
    s = s + e->ledger[j] * (e->mep[j] * (e->ledger[j] < 0. ? 1. : 0.));
  return e->cash + s;
}
double orc_used_margin(const OrcEnv *e) { /* :199-201 */
  double s = fabs(e->ledger[0]) * e->mep[0];
  for (int j = 1; j < e->P.n_assets; ++j) s = s + fabs(e->ledger[j]) * e->mep[j];
  return e->P.required_margin * s;
}
double orc_equity(const OrcEnv *e) { /* :211-213 */
  return e->cash + orc_asset_value(e) - orc_borrowed_margin(e);
}
double orc_borrowed_asset_value(const OrcEnv *e) { /* :219-223 */
  double s = e->ledger[0] * (e->price[0] * (e->ledger[0] < 0. ? 1. : 0.));
  for (int j = 1; j < e->P.n_assets; ++j)
    s = s + e->ledger[j] * (e->price[j] * (e->ledger[j] < 0. ? 1. : 0.));
  return s;
}
double orc_available_margin(const OrcEnv *e) { /* :229-231 */
  return (orc_balance(e) + orc_pnl(e)) / e->P.required_margin;
}
void orc_ledger_normed(const OrcEnv *e, double *out) { /* :140-142 */
  double eq = orc_equity(e);
  for (int j = 0; j < e->P.n_assets; ++j) out[j] = (e->ledger[j] * e->price[j]) / eq;
}
void orc_ledger_normed_full(const OrcEnv *e, double *out) { /* :150-155 */
  double eq = orc_equity(e);
  out[0] = (e->cash - orc_borrowed_margin(e)) / eq;
  for (int j = 0; j < e->P.n_assets; ++j) out[j + 1] = (e->ledger[j] * e->price[j]) / eq;
}
static void abs_norm(double *v, int n) {
  double s = fabs(v[0]);
  for (int j = 1; j < n; ++j) s = s + fabs(v[j]);
  for (int j = 0; j < n; ++j) v[j] = v[j] / s;
}
void orc_ledger_abs_normed(const OrcEnv *e, double *out) { /* :144-148 */
  orc_ledger_normed(e, out);
  abs_norm(out, e->P.n_assets);
}
void orc_ledger_abs_normed_full(const OrcEnv *e, double *out) { /* :164-168 */
  orc_ledger_normed_full(e, out);
  abs_norm(out, e->P.n_assets + 1);
}

int orc_check_risk(const OrcEnv *e) { /* :243-252 */
  double marginRequired = e->P.maintenance_margin * orc_pnl(e);
  if (orc_equity(e) <= -marginRequired) return MDG_RISK_MARGIN_CALL;
  if ((orc_balance(e) + orc_pnl(e)) <= -marginRequired) return MDG_RISK_MARGIN_CALL;
  return MDG_RISK_GREEN;
}

int orc_check_risk_asset(const OrcEnv *e, int i, double units) { /* :254-279 */
  double cashAmount = e->price[i] * units;
  double currentUnits = e->ledger[i];
  if (signbit(units) != signbit(currentUnits)) {
    if (units > -1 * currentUnits) {
      double excess = units + currentUnits;
      if (orc_available_margin(e) <= fabs(e->price[i] * excess) || orc_balance(e) <= 0.)
        return MDG_RISK_INSUFF_MARGIN;
    }
    return MDG_RISK_GREEN;
  } else {
    if (orc_check_risk(e) == MDG_RISK_MARGIN_CALL) return MDG_RISK_MARGIN_CALL;
    if (orc_available_margin(e) <= fabs(cashAmount) || orc_balance(e) <= 0.)
      return MDG_RISK_INSUFF_MARGIN;
    return MDG_RISK_GREEN;
  }
}

void orc_handle_transaction(OrcEnv *e, int i, double transactionPrice, double units,
                            double transactionCost) { /* :284-323 */
  double *currentUnits = &e->ledger[i];
  double *meanEntryPrice = &e->mep[i];
  if (signbit(*currentUnits) != signbit(units)) {
    if (fabs(units) > fabs(*currentUnits)) {
      units += *currentUnits;
      e->cash += *currentUnits * transactionPrice;
      *currentUnits = 0.;
      *meanEntryPrice = transactionPrice;
    }
  } else {
    *meanEntryPrice += (transactionPrice - *meanEntryPrice) * (units / (units + *currentUnits));
  }
  double amount = transactionPrice * units;
  double marginToUse = amount * e->P.required_margin;
  double marginToBorrow = amount - marginToUse;
  double *bm = &e->bm[i];
  *bm += marginToBorrow;
  e->cash -= (marginToUse + transactionCost);
  *currentUnits += units;
  if (fabs(*currentUnits) < 0.000001) {
    *meanEntryPrice = 0.;
    if (*bm > 0.) {
      e->cash -= *bm;
      *bm = 0.;
    }
  }
  if (*bm < 0.) {
    e->cash -= *bm;
    *bm = 0.;
  }
}

/* Broker::handleTransaction(port, i, units)  environments/cpp/Broker.cpp:124-142,171-178 */
int orc_broker_transaction(OrcEnv *e, int i, double units, double *tp, double *tu, double *tc) {
  *tp = 0.; *tu = 0.; *tc = 0.;
  if (units != 0.) {
    int risk = orc_check_risk_asset(e, i, units);
    if (risk == MDG_RISK_GREEN) {
      double currentPrice = e->price[i];
      double slippage = (currentPrice * e->P.slippage_rel) + e->P.slippage_abs;
      double transactionPrice = units < 0 ? (currentPrice - slippage) : (currentPrice + slippage);
      double transactionCost = fabs(units * currentPrice) * e->P.tcost_rel + e->P.tcost_abs;
      orc_handle_transaction(e, i, transactionPrice, units, transactionCost);
      *tp = transactionPrice; *tu = units; *tc = transactionCost;
    }
    return risk;
  }
  return MDG_RISK_GREEN;
}

/* ------------------------------------------------------------------ */
/* Reward shapers (utils/buffers/nstep_buffer.py)                       */
/* ------------------------------------------------------------------ */
static const double EPS32 = 1.1920928955078125e-07; /* np.finfo(np.float32).eps, :20 */
static inline double clip(double x, double lo, double hi) { /* np.clip: NaN propagates */
  if (x != x) return x;
  return x < lo ? lo : (x > hi ? hi : x);
}

static double dsr_value(double A, double B, double r) { /* :80-85 */
  double dA = r - A, dB = r * r - B;
  double v = B - A * A;
  return (B * dA - (A * dB) / 2) / (pow(v * v, 0.75) + EPS32);
}
static double ddr_value(double A, double B, double r) { /* :146-156 */
  if (r > 0.) return (r - A / 2) / (sqrt(B) + EPS32);
  return (B * (r - A / 2) - (A * (r * r)) / 2) / (pow(B, 1.5) + EPS32);
}

/* shaped reward over the current n-step buffer, component c; updates DSR/DDR state */
static double shaper_pop_value(OrcEnv *e, int c) {
  const MdgReward *R = &e->R;
  int n = e->ring_len;
  const double *disc = R->discounts; /* [math.pow(discount, i)], :330, computed by the host */
  switch (R->shaper) {
    case MDG_SHAPER_SUM:
    case MDG_SHAPER_COSINE: { /* :23-27, :182-204 (cosine term already folded in at add time) */
      double s = 0.;
      for (int j = 0; j < n; ++j) s = s + disc[j] * e->ring[j][c];
      return s;
    }
    case MDG_SHAPER_DSR: { /* :62-78 */
      double s = disc[0] * dsr_value(e->A[c], e->B[c], e->ring[0][c]);
      for (int j = 1; j < n; ++j) s = s + disc[j] * dsr_value(e->A[c], e->B[c], e->ring[j][c]);
      s = s / n;
      double r0 = e->ring[0][c], dA = r0 - e->A[c], dB = r0 * r0 - e->B[c]; /* :87-91 */
      e->A[c] += R->adaptation_rate * dA;
      e->B[c] += R->adaptation_rate * dB;
      return clip(s, -1., 1.);
    }
    case MDG_SHAPER_DDR: { /* :128-162 */
      double s = disc[0] * ddr_value(e->A[c], e->B[c], e->ring[0][c]);
      for (int j = 1; j < n; ++j) s = s + disc[j] * ddr_value(e->A[c], e->B[c], e->ring[j][c]);
      s = s / n;
      double r0 = e->ring[0][c], dA = r0 - e->A[c];
      double m = r0 < 0. ? r0 : 0.;
      if (r0 != r0) m = r0; /* np.minimum propagates NaN */
      double dB = m * m - e->B[c];
      e->A[c] += R->adaptation_rate * dA;
      e->B[c] += R->adaptation_rate * dB;
      return clip(s, -1., 1.);
    }
    case MDG_SHAPER_SHARPE: { /* :207-239 */
      if (n == 1) {
        double diff = e->ring[0][c] - 0.;
        diff = (diff != 0.) ? diff : 0.;
        return diff / sqrt(diff * diff);
      }
      double sum = 0., ssq = 0.;
      for (int j = 0; j < n; ++j) {
        double dj = (e->ring[j][c] - 0.) * disc[j];
        if (j == 0) { sum = dj; ssq = dj * dj; } else { sum = sum + dj; ssq = ssq + dj * dj; }
      }
      double num = sum / n;
      double denom = sqrt(ssq / (n - 1));
      double out = (denom != 0.) ? num / denom : 0.;
      return clip(.1 * out, -1., 1.);
    }
    case MDG_SHAPER_SORTINO_A: { /* :242-272 */
      double ex = R->sortino_exp;
      if (n == 1) {
        double diff = e->ring[0][c] - 0.;
        double downside = pow(pow(fabs(diff), ex), 1 / ex);
        return clip(0.1 * ((diff != 0.) ? diff / downside : 0.), -1., 1.);
      }
      double sum = 0., den = 0.;
      for (int j = 0; j < n; ++j) {
        double dj = (e->ring[j][c] - 0.) * disc[j];
        double down = dj < 0. ? dj : 0.;
        if (dj != dj) down = dj;
        if (down < -1.) down = -1.; /* np.clip(.., -1, None) */
        double t = pow(pow(fabs(down), ex) / (n - 1), 1 / ex);
        if (j == 0) { sum = dj; den = t; } else { sum = sum + dj; den = den + t; }
      }
      double num = sum / n;
      double zero_case = (num == 0.) ? 0. : 1.;
      double normal = clip(.1 * (num / den), -1., 1.);
      return (den != 0.) ? normal : zero_case;
    }
    case MDG_SHAPER_SORTINO_B: { /* :276-312 */
      double ex = R->sortino_exp;
      if (n == 1) {
        double diff = e->ring[0][c] - 0.;
        if (diff < -1.) diff = -1.;
        if (diff < 0.) diff = -pow(-diff, 1 / ex);
        return clip(diff, -1., 1.);
      }
      double s = 0.;
      for (int j = 0; j < n; ++j) {
        double dj = (e->ring[j][c] - 0.) * disc[j];
        if (dj < -1.) dj = -1.;
        if (dj < 0.) dj = -pow(-dj, 1 / ex);
        s = (j == 0) ? dj : s + dj;
      }
      return clip(s, -1., 1.);
    }
  }
  return 0.;
}

static void shaper_pop(OrcEnv *e, int ra, OrcStepOut *out) {
  for (int c = 0; c < ra; ++c) out->shaped[out->n_popped][c] = shaper_pop_value(e, c);
  out->n_popped += 1;
  for (int j = 1; j < e->ring_len; ++j) /* self._buffer.pop(0), :352 */
    memcpy(e->ring[j - 1], e->ring[j], sizeof(double) * MDG_MAX_ASSETS);
  e->ring_len -= 1;
}

/* cosine_similarity(next_state.portfolio[-1], desired), :173-177 */
static double cosine_sim(const double *p, const double *q, int n) {
  double pp = p[0] * p[0], qq = q[0] * q[0], pq = p[0] * q[0];
  for (int j = 1; j < n; ++j) {
    pp = pp + p[j] * p[j];
    qq = qq + q[j] * q[j];
    pq = pq + p[j] * q[j];
  }
  return pq / (sqrt(pp) * sqrt(qq));
}

/* ReplayBuffer.add (utils/buffers/replay_buffer.py:68-80) + NStepBuffer (:336-361) */
static void shaper_add(OrcEnv *e, const double *raw, int ra, const double *port, int done,
                       OrcStepOut *out) {
  const MdgReward *R = &e->R;
  double extra = 0.;
  if (R->shaper == MDG_SHAPER_COSINE)
    extra = R->cosine_temp * cosine_sim(port, R->desired_portfolio, e->P.n_assets + 1);
  for (int c = 0; c < ra; ++c)
    e->ring[e->ring_len][c] = (R->shaper == MDG_SHAPER_COSINE) ? raw[c] + extra : raw[c];
  e->ring_len += 1;
  out->n_popped = 0;
  if (e->ring_len >= R->nstep) shaper_pop(e, ra, out);
  if (done)
    while (e->ring_len > 0) shaper_pop(e, ra, out);
}

/* test hook: feed one (raw reward vector, next_state.portfolio[-1], done) to the n-step shaper */
void orc_shaper_feed(OrcEnv *e, const double *raw, int ra, const double *port, int done, OrcStepOut *out) {
  memset(out, 0, sizeof(*out));
  shaper_add(e, raw, ra, port, done, out);
}

/* ------------------------------------------------------------------ */
/* Env  (environments/cpp/Env.h)                                        */
/* ------------------------------------------------------------------ */
void orc_set_ctor_uniforms(OrcEnv *e, const double *u, int n) {
  e->n_ctor_u = (u && n > 0) ? (n > ORC_MAX_CTOR_U ? ORC_MAX_CTOR_U : n) : 0;
  for (int i = 0; i < e->n_ctor_u; ++i) e->ctor_u[i] = u[i];
}

void orc_init(OrcEnv *e, const MdgParams *P, const MdgReward *R, uint64_t seed, int64_t gid) {
  orc_init_inject(e, P, R, seed, gid, 0, 0);
}

void orc_init_inject(OrcEnv *e, const MdgParams *P, const MdgReward *R, uint64_t seed, int64_t gid,
                     const double *ctor_u, int n_ctor_u) {
  memset(e, 0, sizeof(*e));
  if (P->n_gstate > ORC_MAX_GSTATE) {
    fprintf(stderr, "mdg_oracle: n_gstate %d > ORC_MAX_GSTATE %d\n", P->n_gstate, ORC_MAX_GSTATE);
    abort();
  }
  orc_set_ctor_uniforms(e, ctor_u, n_ctor_u);
  e->P = *P;
  if (R) e->R = *R; else { e->R.shaper = MDG_SHAPER_OFF; e->R.nstep = 1; }
  e->seed = seed;
  e->gid = gid;
  e->cash = P->init_cash; /* Portfolio ctor, Portfolio.cpp:7-12,114-116 */
  e->timestamp = 0;
  gen_start(e);
}

static void fill_state_out(const OrcEnv *e, OrcStepOut *out) {
  for (int j = 0; j < e->P.n_assets; ++j) out->price[j] = e->price[j];
  orc_ledger_normed_full(e, out->port);
  out->timestamp = e->timestamp;
}

void orc_reset(OrcEnv *e, const double *normals, const double *uniforms, OrcStepOut *out) {
  /* Env.h:181-187 -> initAccountants :150-165 */
  gen_reset(e);
  for (int j = 0; j < e->P.n_assets; ++j) { e->ledger[j] = 0.; e->mep[j] = 0.; e->bm[j] = 0.; }
  e->cash = e->P.init_cash;
  orc_tick(e, normals, uniforms);
  if (out) {
    memset(out, 0, sizeof(*out));
    fill_state_out(e, out);
  }
}

void orc_step(OrcEnv *e, int mode, const double *units, int asset_idx, const double *normals,
              const double *uniforms, OrcStepOut *out) {
  const int nA = e->P.n_assets;
  memset(out, 0, sizeof(*out));
  double prevEq = orc_equity(e); /* Env.h:190,208,234 */
  double prevVal[MDG_MAX_ASSETS];
  for (int j = 0; j < nA; ++j) prevVal[j] = e->ledger[j] * e->price[j]; /* offpolicy_q.py:141 */
  if (mode == MDG_MODE_MULTI) { /* Broker.cpp:144-158 */
    for (int i = 0; i < nA; ++i)
      out->risk[i] = (uint8_t)orc_broker_transaction(e, i, units[i], &out->tp[i], &out->tu[i], &out->tc[i]);
    out->margin_call = (orc_check_risk(e) == MDG_RISK_MARGIN_CALL);
  } else if (mode == MDG_MODE_SINGLE) { /* Broker.cpp:124-142 */
    int i = asset_idx;
    out->risk[i] = (uint8_t)orc_broker_transaction(e, i, units[0], &out->tp[i], &out->tu[i], &out->tc[i]);
    out->margin_call = (orc_check_risk(e) == MDG_RISK_MARGIN_CALL);
  }
  orc_tick(e, normals, uniforms);
  double currentEq = orc_equity(e);
  double clampv = (mode == MDG_MODE_SINGLE) ? 0.01 : 0.3; /* Env.h:193,212,238 */
  out->reward = log(dmax(currentEq / prevEq, clampv));
  int risk = orc_check_risk(e);
  int done = 0;
  if (mode == MDG_MODE_HOLD) { /* Env.h:194-198 */
    done = (risk == MDG_RISK_GREEN) ? 0 : 1;
    if (currentEq < 0.1 * e->P.init_cash) done = 1;
  } else { /* Env.h:214-223, 241-249 */
    for (int i = 0; i < nA; ++i)
      if (out->risk[i] != MDG_RISK_GREEN && out->risk[i] != MDG_RISK_INSUFF_MARGIN) done = 1;
    if (risk != MDG_RISK_GREEN || currentEq < 0.1 * e->P.init_cash) done = 1;
  }
  out->done = (uint8_t)done;
  fill_state_out(e, out);

  if (e->R.shaper != MDG_SHAPER_OFF && mode != MDG_MODE_HOLD) {
    /* agent reward (only transitions the agent loop stores: initialize_history's no-action
     * steps never reach the replay buffer, preprocessor.py:191-194), modelling/algorithm/offpolicy_q.py:152-164 */
    double r[MDG_MAX_ASSETS];
    for (int j = 0; j < nA; ++j) {
      double curVal = e->ledger[j] * e->price[j];
      double marDiff = out->tu[j] * out->tp[j] + out->tc[j];
      double x = (curVal - prevVal[j] - marDiff) / prevEq;
      x += 1;
      r[j] = log((x != x) ? x : ((x < .35) ? .35 : x)); /* np.maximum propagates NaN */
    }
    int ra = nA;
    if (e->R.reduce_rewards) {
      double s = r[0];
      for (int j = 1; j < nA; ++j) s = s + r[j];
      r[0] = s;
      ra = 1;
    }
    for (int c = 0; c < ra; ++c) out->agent_reward[c] = r[c];
    shaper_add(e, r, ra, out->port, done, out);
  }
}

/* ------------------------------------------------------------------ */
/* Batch                                                                */
/* ------------------------------------------------------------------ */
struct OrcBatch {
  int64_t n;
  OrcEnv *envs;
};

OrcBatch *orc_batch_create(int64_t n, const MdgParams *P, const MdgReward *R, uint64_t seed,
                           int64_t env_offset) {
  OrcBatch *b = (OrcBatch *)malloc(sizeof(OrcBatch));
  b->n = n;
  b->envs = (OrcEnv *)malloc(sizeof(OrcEnv) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) orc_init(&b->envs[i], P, R, seed, env_offset + i);
  return b;
}
void orc_batch_destroy(OrcBatch *b) {
  if (!b) return;
  free(b->envs);
  free(b);
}
OrcEnv *orc_batch_env(OrcBatch *b, int64_t i) { return &b->envs[i]; }

void orc_batch_export_state(const OrcBatch *b, const MdgState *s) {
  const int64_t N = b->n;
  for (int64_t i = 0; i < N; ++i) {
    const OrcEnv *e = &b->envs[i];
    const int nA = e->P.n_assets;
    for (int j = 0; j < nA; ++j) {
      if (s->price) s->price[j * N + i] = e->price[j];
      if (s->ledger) s->ledger[j * N + i] = e->ledger[j];
      if (s->mean_entry) s->mean_entry[j * N + i] = e->mep[j];
      if (s->borrowed) s->borrowed[j * N + i] = e->bm[j];
    }
    if (s->cash) s->cash[i] = e->cash;
    if (s->gstate) for (int g = 0; g < e->P.n_gstate; ++g) s->gstate[g * N + i] = e->gstate[g];
    if (s->timestamp) s->timestamp[i] = e->timestamp;
    int ra = e->R.reduce_rewards ? 1 : nA;
    for (int c = 0; c < ra; ++c) {
      if (s->shaper_A) s->shaper_A[c * N + i] = e->A[c];
      if (s->shaper_B) s->shaper_B[c * N + i] = e->B[c];
    }
    if (s->nstep_len) s->nstep_len[i] = e->ring_len;
    /* s->reset_ts / s->folds are caches of the CUDA path; the oracle has no counterpart */
  }
}

static void write_row(const MdgStepIO *io, const MdgLaunch *L, int nA, int64_t N, int64_t i,
                      int slot, const OrcStepOut *o) {
  for (int j = 0; j < nA; ++j) io->obs_price[((int64_t)slot * nA + j) * N + i] = o->price[j];
  for (int j = 0; j < nA + 1; ++j) io->obs_port[((int64_t)slot * (nA + 1) + j) * N + i] = o->port[j];
  (void)L;
}

/* DQN.action_to_transaction, modelling/algorithm/dqn.py:160-179: units from discrete actions */
void orc_action_units(const OrcEnv *e, const int8_t *actions, int action_atoms, double unit_size, double *units) {
  const double scale = unit_size * orc_available_margin(e); /* :164 self.unit_size * self._env.availableMargin */
  const int half = action_atoms / 2;                        /* :166 action_atoms // 2 */
  for (int j = 0; j < e->P.n_assets; ++j) {
    const double per = scale / e->price[j];                 /* :164-165 ... / self._env.currentPrices */
    units[j] = (double)(actions[j] - half) * per;           /* :169 actions_centered * units */
    if (actions[j] == 0)                                    /* :171-177 action 0 closes an open position */
      units[j] = (e->ledger[j] != 0.) ? -e->ledger[j] : 0.;
  }
}

/* DDPG.action_to_transaction, modelling/algorithm/ddpg.py:182-207: units from target weights (cash first).
 * desired = w / w.sum() in fp32 (torch on the actor's fp32 output; :193-196), widened by the subtraction with the
 * fp64 ledgerNormedFull (:198-201); units = amounts[1:] / currentPrices (:205). */
void orc_weight_units(const OrcEnv *e, const float *w, double *units) {
  const int nA = e->P.n_assets;
  double cur[MDG_MAX_ASSETS + 1];
  float s = w[0];
  for (int j = 1; j <= nA; ++j) s = s + w[j];
  orc_ledger_normed_full(e, cur);
  const double eq = orc_equity(e);
  for (int j = 0; j < nA; ++j) {
    const float d = (s == 0.f) ? w[j + 1] : w[j + 1] / s;
    const double amount = ((double)d - cur[j + 1]) * eq;
    units[j] = amount / e->price[j];
  }
}

void orc_batch_step(OrcBatch *b, const MdgStepIO *io, const MdgLaunch *L, int threads) {
  const int64_t N = b->n;
  (void)threads;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t i = 0; i < N; ++i) {
    OrcEnv *e = &b->envs[i];
    const int nA = e->P.n_assets;
    double nz[16 * MDG_MAX_SINE_COMPONENTS + 64], uz[256], un[MDG_MAX_ASSETS];
    const double *pn = 0, *pu = 0;
    if (io->normals) { for (int s = 0; s < e->P.n_normals; ++s) nz[s] = io->normals[s * N + i]; pn = nz; }
    if (io->uniforms) { for (int s = 0; s < e->P.n_uniforms; ++s) uz[s] = io->uniforms[s * N + i]; pu = uz; }
    if (L->mode == MDG_MODE_MULTI && io->weights)
      orc_weight_units(e, io->weights + i * (nA + 1), un);
    else if (L->mode == MDG_MODE_MULTI && io->actions)
      orc_action_units(e, io->actions + i * nA, L->action_atoms, L->unit_size, un);
    else if (L->mode == MDG_MODE_MULTI) for (int j = 0; j < nA; ++j) un[j] = io->units[i * nA + j];
    if (L->mode == MDG_MODE_SINGLE) un[0] = io->units[i];
    OrcStepOut o;
    orc_step(e, L->mode, un, L->asset_idx, pn, pu, &o);
    write_row(io, L, nA, N, i, L->head, &o);
    io->reward[i] = o.reward;
    io->done[i] = o.done;
    io->margin_call[i] = o.margin_call;
    for (int j = 0; j < nA; ++j) {
      io->trans_price[j * N + i] = o.tp[j];
      io->trans_units[j * N + i] = o.tu[j];
      io->trans_cost[j * N + i] = o.tc[j];
      io->risk[j * N + i] = o.risk[j];
    }
    if (e->R.shaper != MDG_SHAPER_OFF && L->mode != MDG_MODE_HOLD) {
      int ra = e->R.reduce_rewards ? 1 : nA;
      if (io->agent_reward) for (int c = 0; c < ra; ++c) io->agent_reward[c * N + i] = o.agent_reward[c];
      if (io->n_popped) io->n_popped[i] = o.n_popped;
      if (io->shaped_reward)
        for (int k = 0; k < o.n_popped; ++k)
          for (int c = 0; c < ra; ++c) io->shaped_reward[((int64_t)k * ra + c) * N + i] = o.shaped[k][c];
    }
  }
}

void orc_batch_reset(OrcBatch *b, const MdgStepIO *io, const MdgLaunch *L, const uint8_t *mask,
                     int fill_ticks, int clear_nstep, int threads) {
  const int64_t N = b->n;
  (void)threads;
  if (fill_ticks < 1) fill_ticks = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t i = 0; i < N; ++i) {
    if (mask && !mask[i]) continue;
    OrcEnv *e = &b->envs[i];
    const int nA = e->P.n_assets, nN = e->P.n_normals, nU = e->P.n_uniforms;
    if (clear_nstep) e->ring_len = 0; /* offpolicy_q.py:94 */
    for (int t = 0; t < fill_ticks; ++t) {
      double nz[16 * MDG_MAX_SINE_COMPONENTS + 64], uz[256];
      const double *pn = 0, *pu = 0;
      if (io->normals) { for (int s = 0; s < nN; ++s) nz[s] = io->normals[((int64_t)t * nN + s) * N + i]; pn = nz; }
      if (io->uniforms) { for (int s = 0; s < nU; ++s) uz[s] = io->uniforms[((int64_t)t * nU + s) * N + i]; pu = uz; }
      OrcStepOut o;
      if (t == 0) orc_reset(e, pn, pu, &o);
      else orc_step(e, MDG_MODE_HOLD, 0, 0, pn, pu, &o); /* preprocessor.py:191-194 */
      int slot = ((L->head - (fill_ticks - 1 - t)) % L->window + L->window) % L->window;
      write_row(io, L, nA, N, i, slot, &o);
    }
  }
}

void orc_batch_derived(const OrcBatch *b, const MdgDerived *d) {
  const int64_t N = b->n;
  for (int64_t i = 0; i < N; ++i) {
    const OrcEnv *e = &b->envs[i];
    const int nA = e->P.n_assets;
    double v[MDG_MAX_ASSETS + 1];
    if (d->equity) d->equity[i] = orc_equity(e);
    if (d->asset_value) d->asset_value[i] = orc_asset_value(e);
    if (d->pnl) d->pnl[i] = orc_pnl(e);
    if (d->balance) d->balance[i] = orc_balance(e);
    if (d->available_margin) d->available_margin[i] = orc_available_margin(e);
    if (d->used_margin) d->used_margin[i] = orc_used_margin(e);
    if (d->borrowed_margin) d->borrowed_margin[i] = orc_borrowed_margin(e);
    if (d->borrowed_asset_value) d->borrowed_asset_value[i] = orc_borrowed_asset_value(e);
    if (d->risk) d->risk[i] = (uint8_t)orc_check_risk(e);
    for (int j = 0; j < nA; ++j) {
      if (d->position_values) d->position_values[j * N + i] = e->ledger[j] * e->price[j]; /* :170-172 */
      if (d->pnl_positions) /* :188-190 */
        d->pnl_positions[j * N + i] = e->ledger[j] * e->price[j] - e->mep[j] * e->ledger[j];
    }
    if (d->ledger_normed) { orc_ledger_normed(e, v); for (int j = 0; j < nA; ++j) d->ledger_normed[j * N + i] = v[j]; }
    if (d->ledger_abs_normed) { orc_ledger_abs_normed(e, v); for (int j = 0; j < nA; ++j) d->ledger_abs_normed[j * N + i] = v[j]; }
    if (d->ledger_normed_full) { orc_ledger_normed_full(e, v); for (int j = 0; j <= nA; ++j) d->ledger_normed_full[j * N + i] = v[j]; }
    if (d->ledger_abs_normed_full) { orc_ledger_abs_normed_full(e, v); for (int j = 0; j <= nA; ++j) d->ledger_abs_normed_full[j * N + i] = v[j]; }
    if (d->position_values_full) { /* :174-178 */
      d->position_values_full[i] = e->cash - orc_borrowed_margin(e);
      for (int j = 0; j < nA; ++j) d->position_values_full[(j + 1) * N + i] = e->ledger[j] * e->price[j];
    }
    if (d->ledger_full) { /* :157-161 */
      d->ledger_full[i] = e->cash - orc_borrowed_margin(e);
      for (int j = 0; j < nA; ++j) d->ledger_full[(j + 1) * N + i] = e->ledger[j];
    }
  }
}
