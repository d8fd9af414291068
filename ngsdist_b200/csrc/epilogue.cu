// K4 epilogue: deterministic reduction of the K-split partial tiles + the tail of gen_dist (ngsDist.cpp:372-401):
//   cnt <- tot_sites override; d = num / (double) cnt; model 0: d, 1: -log(1-d), 2: -log(1-(d*4/3))*3/4;
// symmetric write of both triangles as gen_dist_slave does (ngsDist.cpp:408-412), 0.0 diagonal (ngsDist.cpp:200).
#include "ngsd_internal.h"

namespace {

struct EpiArgs {
  const double *partials;     // [n_splits][n_tiles][16384] fragment order (dist_dmma.cu)
  const ngsd_tile *tiles;
  const uint32_t *cnt;        // [n_pad][n_pad] or nullptr
  double *out, *num;          // [n_ind][n_ind]
  uint64_t *cntout;
  uint64_t n_ind, n_pad, const_cnt, tot_sites;
  uint32_t n_splits, n_tiles;
  int evol_model;
};

// grid (n_tiles, 8); block 256: thread handles 8 double2 slots of one consumer-warp's fragment block
__global__ void __launch_bounds__(256) k_epilogue(EpiArgs a) {
  const uint32_t t = blockIdx.x;
  const ngsd_tile tl = a.tiles[t];
  const int warp = blockIdx.y;                 // consumer warp whose 64x32 sub-tile this block finishes
  const int wm = warp >> 2, wn = warp & 3;
  const int lane = threadIdx.x & 31;
  for (int frag = threadIdx.x >> 5; frag < 32; frag += 8) {
    const uint64_t e = ((uint64_t) (warp * 32 + frag) * 32 + lane);   // double2 index inside the tile
    double s0 = 0, s1 = 0;
    for (uint32_t q = 0; q < a.n_splits; q++) {                        // fixed order: deterministic
      const double2 v = reinterpret_cast<const double2 *>(a.partials + ((uint64_t) q * a.n_tiles + t) * NGSD_TILE_ELEMS)[e];
      s0 += v.x;
      s1 += v.y;
    }
    const int mi = frag >> 2, ni = frag & 3;
    const uint64_t i = (uint64_t) tl.ti * NGSD_TILE + wm * 64 + mi * 8 + (lane >> 2);
    const uint64_t j0 = (uint64_t) tl.tj * NGSD_TILE + wn * 32 + ni * 8 + 2 * (lane & 3);
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const uint64_t j = j0 + c;
      if (i >= j || j >= a.n_ind) continue;
      const double num = c ? s1 : s0;
      uint64_t cnt = a.cnt ? (uint64_t) a.cnt[i * a.n_pad + j] : a.const_cnt;
      if (a.num) a.num[i * a.n_ind + j] = a.num[j * a.n_ind + i] = num;
      if (a.cntout) a.cntout[i * a.n_ind + j] = a.cntout[j * a.n_ind + i] = cnt;
      if (a.tot_sites > 0) cnt = a.tot_sites;
      double d = num / (double) cnt;
      if (a.evol_model == 1) d = -log(1 - d);
      else if (a.evol_model == 2) d = -log(1 - (d * 4 / 3)) * 3 / 4;
      a.out[i * a.n_ind + j] = a.out[j * a.n_ind + i] = d;
    }
  }
}

__global__ void k_zero_diag(double *out, double *num, uint64_t *cnt, uint64_t n) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i * n + i] = 0.0;
  if (num) num[i * n + i] = 0.0;
  if (cnt) cnt[i * n + i] = 0;
}

}  // namespace

cudaError_t ngsd_launch_epilogue(ngsd_ctx *ctx, const ngsd_epilogue_args &e) {
  EpiArgs a;
  a.partials = ctx->d_partials;
  a.tiles = ctx->d_tiles;
  a.cnt = e.use_cnt ? ctx->d_cnt : nullptr;
  a.out = ctx->d_out;
  a.num = ctx->d_num;
  a.cntout = ctx->d_cntout;
  a.n_ind = ctx->n_ind;
  a.n_pad = ctx->n_pad;
  a.const_cnt = e.const_cnt;
  a.tot_sites = ctx->cfg.tot_sites;
  a.n_splits = e.n_splits;
  a.n_tiles = ctx->n_tiles;
  a.evol_model = ctx->cfg.evol_model;
  k_zero_diag<<<(unsigned) ((ctx->n_ind + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_out, ctx->d_num, ctx->d_cntout, ctx->n_ind);
  k_epilogue<<<dim3(ctx->n_tiles, 8), 256, 0, ctx->stream>>>(a);
  return cudaGetLastError();
}
