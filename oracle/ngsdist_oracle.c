/* ngsdist_oracle.c -- CPU restatement of ngsDist's pairwise-distance hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ngsdist_b200/ may include, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs use it, and only as the checker.
 *
 * Parity status: PINNED against the unmodified reference binary built here from
 * /root/reference (oracle/Makefile `ref` -> oracle/_ref/ngsDist) over the flag matrix in
 * tests/golden/ (script: tests/golden/make_golden.py), and against the GSL gsl_rng_taus
 * known answer (seed 1 -> 10 000th draw 2733957125).  The reference's own md5 goldens
 * (examples/test.md5) are unreachable: their inputs are not shipped.
 *
 * Each function cites the reference lines it restates.  Plain C11, compiled with
 * -ffp-contract=off so that no FMA contraction changes the reference's rounding.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define N_GENO 3
#define NGSD_INF 1e15      /* shared/gen_func.hpp:15 */
#define NGSD_EPS 1e-5      /* shared/gen_func.hpp:16 */

/* ------------------------------------------------------------------ RNG ---
 * gsl_rng_taus semantics (GSL rng/taus.c; reference call sites ngsDist.cpp:179-180,
 * shared/gen_func.cpp:117-119).  state = {s1,s2,s3}. */
static uint32_t taus_step(uint32_t *st) {
#define TAUS(s, a, b, c, d) ((((s) & (c)) << (d)) ^ ((((s) << (a)) ^ (s)) >> (b)))
  st[0] = TAUS(st[0], 13, 19, 4294967294u, 12);
  st[1] = TAUS(st[1], 2, 25, 4294967288u, 4);
  st[2] = TAUS(st[2], 3, 11, 4294967280u, 17);
#undef TAUS
  return st[0] ^ st[1] ^ st[2];
}

void ngsd_oracle_taus_set(uint32_t *st, uint32_t seed) {
  if (seed == 0) seed = 1;
  st[0] = 69069u * seed;
  st[1] = 69069u * st[0];
  st[2] = 69069u * st[1];
  for (int i = 0; i < 6; i++) taus_step(st);
}

uint32_t ngsd_oracle_taus_get(uint32_t *st) { return taus_step(st); }

/* rnd_map_data (ngsDist.cpp:416-437) as a site map: map[block*bs+s] = rnd_block*bs+s,
 * rnd_block = floor(0 + uniform*n_blocks) (gen_func.cpp:117-119).  Consumes n_blocks draws. */
void ngsd_oracle_boot_map(uint32_t *st, uint64_t n_blocks, uint64_t block_size, uint64_t *site_map) {
  for (uint64_t b = 0; b < n_blocks; b++) {
    double u = taus_step(st) / 4294967296.0;
    uint64_t rb = (uint64_t) floor(0 + u * (double) (n_blocks - 0));
    for (uint64_t s = 0; s < block_size; s++) site_map[b * block_size + s] = rb * block_size + s;
  }
}

/* ------------------------------------------------------------ front end --- */
/* logsum (gen_func.cpp:135-151) */
static double logsum3(const double *a) {
  double M = a[0];
  for (int i = 1; i < N_GENO; i++) M = (a[i] >= M ? a[i] : M);   /* max() macro, gen_func.hpp:22 */
  if (M == -INFINITY) return -INFINITY;
  double sum = 0;
  for (int i = 0; i < N_GENO; i++) sum += exp(a[i] - M);
  return log(sum) + M;
}

/* post_prob with prior == NULL (gen_func.cpp:920-932) */
static void post_prob3(double *pp) {
  double norm = logsum3(pp);
  for (int g = 0; g < N_GENO; g++) pp[g] -= norm;
}

/* call_geno with log_scale = true, miss_data = 0 (gen_func.cpp:886-914; array_max_pos/min_pos :73-98) */
static void call_geno3(double *geno, double N_thresh, double call_thresh) {
  int max_pos = 0, min_pos = 0;
  double mx = -INFINITY, mn = +INFINITY;
  for (int g = 0; g < N_GENO; g++) if (geno[g] > mx) { max_pos = g; mx = geno[g]; }
  for (int g = 0; g < N_GENO; g++) if (geno[g] < mn) { min_pos = g; mn = geno[g]; }
  double max_pp = exp(geno[max_pos]);
  if (geno[min_pos] == geno[max_pos]) max_pp = -1;
  if (max_pp < N_thresh)
    for (int g = 0; g < N_GENO; g++) geno[g] = log((double) 1 / N_GENO);
  if (max_pp >= call_thresh) {
    for (int g = 0; g < N_GENO; g++) geno[g] = -NGSD_INF;
    geno[max_pos] = log(1);
  }
}

/* Front end of the hot path for one raw triple.
 *   kind 0: binary reader path (read_data.cpp:29-47): log unless in_log, -inf -> -1e15
 *           (conv_space, gen_func.cpp:123-130), normalise, NaN -> error (returns 1).
 *   kind 1: text --probs path (read_data.cpp:83-87,98): log unless in_log, NO clamp, normalise.
 *   then main's loop (ngsDist.cpp:165-174): optional call_geno in log space, exp() with
 *   conv_space's -inf clamp (exp never yields -inf, so the clamp is inert). */
static int frontend_one(const double *x, int kind, int in_log, int do_call, double N_thresh, double call_thresh, double *p) {
  double L[N_GENO];
  for (int g = 0; g < N_GENO; g++) {
    L[g] = in_log ? x[g] : log(x[g]);
    if (kind == 0 && !in_log && L[g] == -INFINITY) L[g] = -NGSD_INF;
  }
  post_prob3(L);
  if (kind == 0 && (isnan(L[0]) || isnan(L[1]) || isnan(L[2]))) return 1;
  if (do_call) call_geno3(L, N_thresh, call_thresh);
  for (int g = 0; g < N_GENO; g++) p[g] = exp(L[g]);
  return 0;
}

/* raw: [site][ind][3] (binary layout, read_data.cpp:28-31).  P: [ind][site][3], normal space.
 * Returns 0, 1 (NaN found) or 2 (N_thresh > call_thresh, gen_func.cpp:887-888). */
int ngsd_oracle_frontend(const double *raw, uint64_t n_ind, uint64_t n_sites, int kind, int in_log, int do_call,
                         double N_thresh, double call_thresh, double *P) {
  if (do_call && N_thresh > call_thresh) return 2;
  int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
  for (uint64_t s = 0; s < n_sites; s++)
    for (uint64_t i = 0; i < n_ind; i++)
      bad |= frontend_one(raw + (s * n_ind + i) * N_GENO, kind, in_log, do_call, N_thresh, call_thresh,
                          P + (i * n_sites + s) * N_GENO);
  return bad;
}

/* Genotype (non --probs) text input (read_data.cpp:88-95,98): g in {0,1,2} -> that entry log(1)=0,
 * the others keep init_ptr's -INF fill (read_data.cpp:21, "-INF" = -1e15); g < 0 -> log(1/3) x3;
 * g > 2 -> error (returns 3).  Then post_prob and main's exp(). */
int ngsd_oracle_frontend_geno(const int32_t *codes /* [site][ind] */, uint64_t n_ind, uint64_t n_sites, double *P) {
  for (uint64_t s = 0; s < n_sites; s++)
    for (uint64_t i = 0; i < n_ind; i++) {
      int g = codes[s * n_ind + i];
      double L[N_GENO] = { -NGSD_INF, -NGSD_INF, -NGSD_INF };
      if (g >= 0) {
        if (g > 2) return 3;
        L[g] = log(1);
      } else
        L[0] = L[1] = L[2] = log((double) 1 / N_GENO);
      post_prob3(L);
      double *p = P + (i * n_sites + s) * N_GENO;
      for (int k = 0; k < N_GENO; k++) p[k] = exp(L[k]);
    }
  return 0;
}

/* miss_data (gen_func.cpp:862-868; abs macro gen_func.hpp:20) */
static int miss_data3(const double *g) {
  double d01 = g[0] - g[1], d12 = g[1] - g[2];
  d01 = d01 >= 0 ? d01 : -d01;
  d12 = d12 >= 0 ? d12 : -d12;
  return d01 < NGSD_EPS && d12 < NGSD_EPS;
}

void ngsd_oracle_miss_mask(const double *P, uint64_t n_ind, uint64_t n_sites, uint8_t *present /* [ind][site] */) {
  for (uint64_t k = 0; k < n_ind * n_sites; k++) present[k] = !miss_data3(P + k * N_GENO);
}

/* ------------------------------------------------------------- em2 -------
 * emOptim2.cpp:69-135 specialised to one site (GL1->x == 1), dim 9. */
static void normalize9(double *t) {
  double s = 0;
  for (int i = 0; i < 9; i++) s += t[i];
  for (int i = 0; i < 9; i++) t[i] /= s;
}

static double lik2_one(const double *sfs, const double *a, const double *b) {
  double tmp = 0;
  int inc = 0;
  for (int x = 0; x < 3; x++)
    for (int y = 0; y < 3; y++) tmp += sfs[inc++] * a[x] * b[y];
  return 0 + log(tmp);
}

static void emstep2_one(const double *pre, const double *a, const double *b, double *post) {
  double inner[9];
  for (int x = 0; x < 9; x++) post[x] = 0.0;
  int inc = 0;
  for (int x = 0; x < 3; x++)
    for (int y = 0; y < 3; y++) { inner[inc] = pre[inc] * a[x] * b[y]; inc++; }
  normalize9(inner);
  for (int x = 0; x < 9; x++) post[x] += inner[x];
  normalize9(post);
}

/* returns number of EM iterations executed (for T-agreement statistics) */
int ngsd_oracle_em2(double *sfs, const double *a, const double *b, double tole, int maxIter) {
  double oldLik = lik2_one(sfs, a, b), lik;
  double tmp[9];
  int it;
  for (it = 0; it < maxIter; it++) {
    emstep2_one(sfs, a, b, tmp);
    for (int i = 0; i < 9; i++) sfs[i] = tmp[i];
    lik = lik2_one(sfs, a, b);
    if (fabs(lik - oldLik) < tole) { oldLik = lik; it++; break; }
    oldLik = lik;
  }
  return it;
}

/* ------------------------------------------------------------ gen_dist ---
 * ngsDist.cpp:325-404 for one pair.  site_map == NULL -> identity (replicate 0); otherwise
 * geno_lkl[i][s] = in_geno_lkl[i][site_map[s]] (the re-pointing done by rnd_map_data).
 * P is [ind][n_sites_total][3].  Returns the model-transformed distance; *num_out, *cnt_out
 * receive the raw accumulator and the valid-site count (the --verbose 3 values, :366-367). */
static double gen_dist_pair(const double *P, uint64_t n_sites_total, const uint64_t *site_map, uint64_t n_eff,
                            const double *score, int indep, int pairwise_del, uint64_t tot_sites, int evol_model,
                            uint64_t i1, uint64_t i2, double *num_out, uint64_t *cnt_out, uint64_t *iters_out) {
  uint64_t cnt = 0, iters = 0;
  double dist = 0;
  for (uint64_t s = 0; s < n_eff; s++) {
    uint64_t src = site_map ? site_map[s] : s;
    const double *a = P + (i1 * n_sites_total + src) * N_GENO;
    const double *b = P + (i2 * n_sites_total + src) * N_GENO;
    if (pairwise_del && (miss_data3(a) || miss_data3(b))) continue;
    double sfs[9];
    for (int k = 0; k < 9; k++) sfs[k] = (double) 1 / 9;
    if (!indep) iters += ngsd_oracle_em2(sfs, a, b, 0.001, 50);
    for (int g1 = 0; g1 < 3; g1++)
      for (int g2 = 0; g2 < 3; g2++) dist += score[3 * g1 + g2] * (indep ? a[g1] * b[g2] : sfs[3 * g1 + g2]);
    cnt++;
  }
  if (num_out) *num_out = dist;
  if (cnt_out) *cnt_out = cnt;
  if (iters_out) *iters_out = iters;
  if (tot_sites > 0) cnt = tot_sites;
  dist /= (double) cnt;
  if (evol_model == 0) dist = dist;
  else if (evol_model == 1) dist = -log(1 - dist);
  else if (evol_model == 2) dist = -log(1 - (dist * 4 / 3)) * 3 / 4;
  else dist = NAN; /* models 3..6: error() in the reference (ngsDist.cpp:387-401) */
  return dist;
}

/* All pairs (dispatch loop ngsDist.cpp:244-262 + gen_dist_slave :408-412).  dist: n_ind x n_ind
 * row-major, symmetric, 0.0 diagonal (init_ptr at :200).  num / cnt / em_iters optional (may be NULL),
 * same shape, upper and lower triangle both filled.  Returns 0, or 4 for an unsupported model. */
int ngsd_oracle_distances(const double *P, uint64_t n_ind, uint64_t n_sites_total, const uint64_t *site_map,
                          uint64_t n_eff, const double *score, int indep, int pairwise_del, uint64_t tot_sites,
                          int evol_model, double *dist, double *num, uint64_t *cnt, uint64_t *em_iters) {
  if (evol_model < 0 || evol_model > 2) return 4;
  for (uint64_t i = 0; i < n_ind; i++) dist[i * n_ind + i] = 0.0;
  int64_t n_pairs = (int64_t) (n_ind * (n_ind - 1) / 2);
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t c = 0; c < n_pairs; c++) {
    /* unrank c -> (i1 < i2) in the row-major upper-triangle order of the dispatch loop */
    uint64_t i1 = 0, rem = (uint64_t) c;
    while (rem >= n_ind - 1 - i1) { rem -= n_ind - 1 - i1; i1++; }
    uint64_t i2 = i1 + 1 + rem;
    double nm; uint64_t ct, it;
    double d = gen_dist_pair(P, n_sites_total, site_map, n_eff, score, indep, pairwise_del, tot_sites, evol_model,
                             i1, i2, &nm, &ct, &it);
    dist[i1 * n_ind + i2] = dist[i2 * n_ind + i1] = d;
    if (num) num[i1 * n_ind + i2] = num[i2 * n_ind + i1] = nm;
    if (cnt) cnt[i1 * n_ind + i2] = cnt[i2 * n_ind + i1] = ct;
    if (em_iters) em_iters[i1 * n_ind + i2] = em_iters[i2 * n_ind + i1] = it;
  }
  return 0;
}

/* A rectangular sub-block of pairs (rows i in [r0,r1), cols j in [c0,c1), only i<j computed):
 * used for spot checks at shapes where the full matrix would take too long on the CPU. */
int ngsd_oracle_distances_block(const double *P, uint64_t n_ind, uint64_t n_sites_total, const uint64_t *site_map,
                                uint64_t n_eff, const double *score, int indep, int pairwise_del, uint64_t tot_sites,
                                int evol_model, uint64_t r0, uint64_t r1, uint64_t c0, uint64_t c1, double *dist,
                                double *num, uint64_t *cnt) {
  if (evol_model < 0 || evol_model > 2) return 4;
  uint64_t w = c1 - c0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = (int64_t) r0; i < (int64_t) r1; i++)
    for (uint64_t j = c0; j < c1; j++) {
      uint64_t o = ((uint64_t) i - r0) * w + (j - c0);
      if ((uint64_t) i >= j) { dist[o] = 0; if (num) num[o] = 0; if (cnt) cnt[o] = 0; continue; }
      double nm; uint64_t ct;
      dist[o] = gen_dist_pair(P, n_sites_total, site_map, n_eff, score, indep, pairwise_del, tot_sites, evol_model,
                              (uint64_t) i, j, &nm, &ct, NULL);
      if (num) num[o] = nm;
      if (cnt) cnt[o] = ct;
    }
  return 0;
}

/* ----------------------------------------------------- synthetic inputs ---
 * deterministic, transcendental-free, so CPU and GPU generate bit-identical
 * raw values.  h = splitmix64(seed ^ splitmix64((s*n_ind+i)*4+g)); u = ((h>>11)+0.5)*2^-53;
 * x_g = u^8.  Missing when splitmix64((seed+1) ^ (s*n_ind+i)) < miss_rate*2^64 -> (1/3,1/3,1/3). */
static uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

void ngsd_oracle_synth_raw(uint64_t seed, double miss_rate, uint64_t n_ind, uint64_t site0, uint64_t n_sites, double *raw) {
  /* threshold: miss_rate * 2^64, saturating */
  long double t = (long double) miss_rate * 18446744073709551616.0L;
  uint64_t thr = t >= 18446744073709551615.0L ? UINT64_MAX : (uint64_t) t;
#pragma omp parallel for schedule(static)
  for (uint64_t s = 0; s < n_sites; s++)
    for (uint64_t i = 0; i < n_ind; i++) {
      uint64_t idx = (site0 + s) * n_ind + i;
      double *x = raw + (s * n_ind + i) * 3;
      if (miss_rate > 0 && splitmix64((seed + 1) ^ idx) < thr) {
        x[0] = x[1] = x[2] = 1.0 / 3.0;
        continue;
      }
      for (uint64_t g = 0; g < 3; g++) {
        uint64_t h = splitmix64(seed ^ splitmix64(idx * 4 + g));
        double u = ((double) (h >> 11) + 0.5) * 0x1p-53;
        double u2 = u * u, u4 = u2 * u2;
        x[g] = u4 * u4;
      }
    }
}
