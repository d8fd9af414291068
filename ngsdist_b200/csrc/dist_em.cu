// K2b dist_em: the default --probs path of gen_dist, i.e. WITHOUT indep_geno (ngsDist.cpp:340-353): for every
// pair-site the 9-cell joint genotype table is re-estimated by em2() (emOptim2.cpp:112-135; emStep2 :91-109,
// lik2 :77-89, normalize :69-75) from ONE site of data, tol 0.001, <= 50 iterations, and the distance term is
// sum(score o sfs_T).  That is not bilinear in (p_i, p_j), so it cannot be a GEMM; it runs on the FP64
// CUDA cores.
//
// Closed form used here (validated there on 6e6 pair-sites and here against the oracle): with a uniform
// start, after t EM steps sfs_t = (a^t / S_a(t)) (x) (b^t / S_b(t)) with element-wise powers and S_x(t) = sum_g x_g^t,
// and the reference's log-likelihood after step t is log( S(t+1) / S(t) ), S = S_a * S_b, S(0) = 9.  It stops at the
// first T >= 1 with |log r_T - log r_{T-1}| < 0.001  <=>  S(T+1) * S(T-1) < e^0.001 * S(T)^2  (the ratio is >= 1 by
// log-convexity of power sums), else at T = 50.  Powers are taken of a / max(a), b / max(b) so nothing underflows;
// the table and the test are scale invariant.  One thread owns one pair and walks its sites at its own pace: every
// loop trip is ONE EM step of the thread's current site, so lanes with different T do not wait for each other.
#include <stdlib.h>

#include <algorithm>

#include "ngsd_internal.h"

namespace {

constexpr int kGroupChunks = 4;                 // chunks (of 8 sites) staged per shared-memory group
constexpr int kGroupSites = kGroupChunks * NGSD_SC;
constexpr int kRound = 4;                       // EM steps between two finish rounds
constexpr double kExpTol = 1.0010005001667084;  // e^0.001 (tole of ngsDist.cpp:349)

struct EmArgs {
  const double *Apack;          // packed posterior planes (A operand of dist_dmma)
  const double *weights;        // [NC*8] per-site bootstrap weights or nullptr
  const uint32_t *chunk_ids;    // active chunk list or nullptr (identity)
  double *partials;             // [n_splits][n16*16][n16*16]
  uint64_t NC;
  uint32_t n_chunks, n_splits, n16;
  double score[9];
};

// grid (n16 [tj], n16 [ti], n_splits); block 256 = 16 x 16 pairs
template <bool WEIGHTED>
__global__ void __launch_bounds__(256) k_dist_em(EmArgs a) {
  const uint32_t tj = blockIdx.x, ti = blockIdx.y;
  if (tj < ti) return;
  __shared__ double s_row[16 * 3 * kGroupSites];     // [(r*3+g)*64 + s]   alpha = a / max(a); 0 marks "skip"
  __shared__ double s_col[kGroupSites * 3 * 16];     // [(s*3+g)*16 + c]   beta
  __shared__ double s_w[kGroupSites];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const uint64_t i = (uint64_t) ti * 16 + ty, j = (uint64_t) tj * 16 + tx;
  const bool live = i < j;                             // upper triangle only
  const uint32_t c0 = (uint32_t) (((uint64_t) blockIdx.z * a.n_chunks) / a.n_splits);
  const uint32_t c1 = (uint32_t) (((uint64_t) (blockIdx.z + 1) * a.n_chunks) / a.n_splits);
  const double D00 = a.score[0], D01 = a.score[1], D02 = a.score[2], D10 = a.score[3], D11 = a.score[4], D12 = a.score[5],
               D20 = a.score[6], D21 = a.score[7], D22 = a.score[8];
  double acc = 0.0;

  for (uint32_t cg = c0; cg < c1; cg += kGroupChunks) {
    const int nch = (int) min((uint32_t) kGroupChunks, c1 - cg);
    const int ns = nch * NGSD_SC;
    __syncthreads();
    // ---- stage + normalise: 32 individuals (16 rows, 16 cols) x ns sites; thread -> (ind, site) ----
    for (int e = tid; e < 32 * kGroupSites; e += 256) {
      const int s = e % kGroupSites, ind = e / kGroupSites;         // ind 0..15 rows, 16..31 cols
      double v0 = 0, v1 = 0, v2 = 0;
      if (s < ns) {
        const uint64_t chunk = a.chunk_ids ? a.chunk_ids[cg + (s >> 3)] : (uint64_t) (cg + (s >> 3));
        const uint64_t gi = (ind < 16) ? (uint64_t) ti * 16 + ind : (uint64_t) tj * 16 + (ind - 16);
        const uint64_t rb = gi >> 7, r = gi & 127;
        const int q = s & 7;
        const double *src = a.Apack + (rb * a.NC + chunk) * NGSD_TILE_DOUBLES + (uint64_t) ((q >> 2) * 16 + (r >> 3)) * 32 + (r & 7) * 4 + (q & 3);
        v0 = src[0]; v1 = src[2 * 512]; v2 = src[4 * 512];          // planes g = 0,1,2 -> k4-group g*2 + h
        const double m = fmax(v0, fmax(v1, v2));
        if (m > 0) { v0 /= m; v1 /= m; v2 /= m; } else { v0 = v1 = v2 = 0; }   // all-zero = padded or pairwise-deleted
        if (WEIGHTED && ind == 0) s_w[s] = a.weights[chunk * NGSD_SC + q];
      }
      if (ind < 16) {
        s_row[(ind * 3 + 0) * kGroupSites + s] = v0; s_row[(ind * 3 + 1) * kGroupSites + s] = v1; s_row[(ind * 3 + 2) * kGroupSites + s] = v2;
      } else {
        const int c = ind - 16;
        s_col[(s * 3 + 0) * 16 + c] = v0; s_col[(s * 3 + 1) * 16 + c] = v1; s_col[(s * 3 + 2) * 16 + c] = v2;
      }
    }
    __syncthreads();

    // ---- EM over the group's sites.  Lanes walk their sites at their own pace; to keep divergence cheap the warp
    // alternates kRound predicated EM steps (lanes whose site has converged idle) with ONE finish round in which
    // every converged lane adds its term and fetches its next site. ----
    int s = -1, t = 1;
    double a0 = 0, a1 = 0, a2 = 0, b0 = 0, b1 = 0, b2 = 0;     // alpha, beta of the current site
    double p0 = 0, p1 = 0, p2 = 0, q0 = 0, q1 = 0, q2 = 0;     // alpha^t, beta^t
    double Sprev = 9.0, Scur = 1.0;
    bool active = live, done = false;
    auto next_site = [&]() {
      for (;;) {
        if (++s >= ns) { active = false; return; }
        a0 = s_row[(ty * 3 + 0) * kGroupSites + s]; a1 = s_row[(ty * 3 + 1) * kGroupSites + s]; a2 = s_row[(ty * 3 + 2) * kGroupSites + s];
        b0 = s_col[(s * 3 + 0) * 16 + tx]; b1 = s_col[(s * 3 + 1) * 16 + tx]; b2 = s_col[(s * 3 + 2) * 16 + tx];
        const bool valid = (a0 + a1 + a2 > 0) && (b0 + b1 + b2 > 0) && (!WEIGHTED || s_w[s] != 0.0);
        if (valid) break;
      }
      p0 = a0; p1 = a1; p2 = a2; q0 = b0; q1 = b1; q2 = b2;
      t = 1; Sprev = 9.0; Scur = (a0 + a1 + a2) * (b0 + b1 + b2);
    };
    if (active) next_site();
    while (__any_sync(0xffffffffu, active)) {
#pragma unroll
      for (int r = 0; r < kRound; r++) {
        if (active && !done) {
          const double n0 = p0 * a0, n1 = p1 * a1, n2 = p2 * a2, m0 = q0 * b0, m1 = q1 * b1, m2 = q2 * b2;
          const double Snext = (n0 + n1 + n2) * (m0 + m1 + m2);
          if (Snext * Sprev < kExpTol * (Scur * Scur) || t >= 50) {
            done = true;
          } else {
            p0 = n0; p1 = n1; p2 = n2; q0 = m0; q1 = m1; q2 = m2;
            Sprev = Scur; Scur = Snext; t++;
          }
        }
      }
      if (active && done) {
        // sum(score o sfs_T), sfs_T = (p (x) q) / (S_a(T) S_b(T)); row-major (g1, g2) order as ngsDist.cpp:351-353
        double d = D00 * (p0 * q0);
        d += D01 * (p0 * q1); d += D02 * (p0 * q2);
        d += D10 * (p1 * q0); d += D11 * (p1 * q1); d += D12 * (p1 * q2);
        d += D20 * (p2 * q0); d += D21 * (p2 * q1); d += D22 * (p2 * q2);
        d /= Scur;
        acc += WEIGHTED ? s_w[s] * d : d;
        done = false;
        next_site();
      }
    }
  }
  const uint64_t ld = (uint64_t) a.n16 * 16;
  a.partials[((uint64_t) blockIdx.z * ld + i) * ld + j] = live ? acc : 0.0;
}

struct EmEpiArgs {
  const double *partials;
  const uint32_t *cnt;        // [n_pad][n_pad] or nullptr
  double *out, *num;
  uint64_t *cntout;
  uint64_t n_ind, n_pad, ld, const_cnt, tot_sites;
  uint32_t n_splits;
  int evol_model;
};

// tail of gen_dist (ngsDist.cpp:372-401) for the EM path; fixed-order reduction over the K splits
__global__ void __launch_bounds__(256) k_epilogue_em(EmEpiArgs a) {
  const uint64_t j = (uint64_t) blockIdx.x * 16 + (threadIdx.x & 15), i = (uint64_t) blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i >= a.n_ind || j >= a.n_ind) return;
  if (i == j) {
    a.out[i * a.n_ind + i] = 0.0;
    if (a.num) a.num[i * a.n_ind + i] = 0.0;
    if (a.cntout) a.cntout[i * a.n_ind + i] = 0;
    return;
  }
  if (i > j) return;
  double num = 0;
  for (uint32_t q = 0; q < a.n_splits; q++) num += a.partials[((uint64_t) q * a.ld + i) * a.ld + j];
  uint64_t cnt = a.cnt ? (uint64_t) a.cnt[i * a.n_pad + j] : a.const_cnt;
  if (a.num) a.num[i * a.n_ind + j] = a.num[j * a.n_ind + i] = num;
  if (a.cntout) a.cntout[i * a.n_ind + j] = a.cntout[j * a.n_ind + i] = cnt;
  if (a.tot_sites > 0) cnt = a.tot_sites;
  double d = num / (double) cnt;
  if (a.evol_model == 1) d = -log(1 - d);
  else if (a.evol_model == 2) d = -log(1 - (d * 4 / 3)) * 3 / 4;
  a.out[i * a.n_ind + j] = a.out[j * a.n_ind + i] = d;
}

}  // namespace

uint32_t ngsd_em_splits(const ngsd_ctx *ctx, uint32_t n_chunks) {
  const uint64_t n16 = (ctx->n_ind + 15) / 16, tiles = n16 * (n16 + 1) / 2;
  uint64_t want = ((uint64_t) 24 * ctx->n_sm + tiles - 1) / tiles;            // >= ~3 waves of 8 CTAs per SM
  const uint64_t maxs = std::max<uint64_t>(1, n_chunks / kGroupChunks);
  const uint64_t cap_mem = std::max<uint64_t>(1, ((uint64_t) 2 << 30) / (n16 * 16 * n16 * 16 * 8));
  want = std::max<uint64_t>(1, std::min(std::min(want, maxs), cap_mem));
  return (uint32_t) std::min<uint64_t>(want, 65535);
}

cudaError_t ngsd_launch_dist_em(ngsd_ctx *ctx, uint32_t n_chunks, uint32_t n_splits, bool weighted) {
  EmArgs a;
  a.Apack = ctx->Apack;
  a.weights = weighted ? ctx->d_weights : nullptr;
  a.chunk_ids = weighted ? ctx->d_chunk_ids : nullptr;
  a.partials = ctx->d_partials;
  a.NC = ctx->NC;
  a.n_chunks = n_chunks;
  a.n_splits = n_splits;
  a.n16 = (uint32_t) ((ctx->n_ind + 15) / 16);
  for (int k = 0; k < 9; k++) a.score[k] = ctx->cfg.score[k];
  dim3 grid(a.n16, a.n16, n_splits);
  if (weighted)
    k_dist_em<true><<<grid, 256, 0, ctx->stream>>>(a);
  else
    k_dist_em<false><<<grid, 256, 0, ctx->stream>>>(a);
  return cudaGetLastError();
}

cudaError_t ngsd_launch_epilogue_em(ngsd_ctx *ctx, uint32_t n_splits, uint64_t const_cnt, bool use_cnt) {
  EmEpiArgs a;
  a.partials = ctx->d_partials;
  a.cnt = use_cnt ? ctx->d_cnt : nullptr;
  a.out = ctx->d_out;
  a.num = ctx->d_num;
  a.cntout = ctx->d_cntout;
  a.n_ind = ctx->n_ind;
  a.n_pad = ctx->n_pad;
  const uint32_t n16 = (uint32_t) ((ctx->n_ind + 15) / 16);
  a.ld = (uint64_t) n16 * 16;
  a.const_cnt = const_cnt;
  a.tot_sites = ctx->cfg.tot_sites;
  a.n_splits = n_splits;
  a.evol_model = ctx->cfg.evol_model;
  k_epilogue_em<<<dim3(n16, n16), 256, 0, ctx->stream>>>(a);
  return cudaGetLastError();
}
