// K3 mask_count: shared-site counts under --pairwise_del, bit-exact.
//   cnt(i,j) = sum_s w_s * m_i(s) * m_j(s)        m = !miss_data (gen_func.cpp:862-868; skip at ngsDist.cpp:335-338,362)
// as an integer AND+POPC "mask GEMM" over 64-site words (1-bit MMA is emulated on sm_100a, SURVEY App. D, so the
// CUDA-core LOP3+POPC path is the native one).  Bootstrap weights are small integers: the host decomposes them into
// level masks W_v = {s : w_s >= v}, so cnt = sum_v popc(m_i & m_j & W_v); the kernel just walks a list of
// (word index, level-mask word) entries -- replicate 0 is the list of all words with all-ones masks.
#include "ngsd_internal.h"

namespace {

constexpr int kEB = 16;   // entries per shared-memory batch

struct CountArgs {
  const uint64_t *mask;       // [RB][NW][128]
  const uint32_t *ent_word;
  const uint64_t *ent_mask;
  const ngsd_tile *tiles;
  uint32_t *cnt;              // [n_pad][n_pad], zeroed
  uint64_t NW, n_pad, n_entries;
  uint32_t n_splits;
};

// grid (n_tiles, n_splits); block 256 = 16 x 16 threads, each 8 x 8 pairs (rows rr*16+ty, cols cc*16+tx)
__global__ void __launch_bounds__(256) k_mask_count(CountArgs a) {
  __shared__ uint64_t sa[kEB][128];
  __shared__ uint64_t sb[kEB][128];
  const ngsd_tile tl = a.tiles[blockIdx.x];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const uint64_t e0 = (a.n_entries * blockIdx.y) / a.n_splits, e1 = (a.n_entries * (blockIdx.y + 1)) / a.n_splits;
  const uint64_t *ma = a.mask + (uint64_t) tl.ti * a.NW * 128;
  const uint64_t *mb = a.mask + (uint64_t) tl.tj * a.NW * 128;
  uint32_t acc[8][8];
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c < 8; c++) acc[r][c] = 0;

  for (uint64_t eb = e0; eb < e1; eb += kEB) {
    const int nb = (int) ((e1 - eb) < kEB ? (e1 - eb) : kEB);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kEB * 128 / 256; k++) {
      const int idx = k * 256 + tid, e = idx >> 7, r = idx & 127;
      uint64_t va = 0, vb = 0;
      if (e < nb) {
        const uint64_t w = a.ent_word[eb + e];
        va = ma[w * 128 + r];
        vb = mb[w * 128 + r] & a.ent_mask[eb + e];
      }
      sa[e][r] = va;
      sb[e][r] = vb;
    }
    __syncthreads();
#pragma unroll 2
    for (int e = 0; e < kEB; e++) {
      uint64_t av[8], bv[8];
#pragma unroll
      for (int r = 0; r < 8; r++) av[r] = sa[e][r * 16 + ty];
#pragma unroll
      for (int c = 0; c < 8; c++) bv[c] = sb[e][c * 16 + tx];
#pragma unroll
      for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 8; c++) acc[r][c] += __popcll(av[r] & bv[c]);
    }
  }
#pragma unroll
  for (int r = 0; r < 8; r++)
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const uint64_t i = (uint64_t) tl.ti * 128 + r * 16 + ty, j = (uint64_t) tl.tj * 128 + c * 16 + tx;
      if (acc[r][c]) atomicAdd(&a.cnt[i * a.n_pad + j], acc[r][c]);
    }
}

}  // namespace

cudaError_t ngsd_launch_mask_count(ngsd_ctx *ctx, uint64_t n_entries) {
  cudaError_t e = cudaMemsetAsync(ctx->d_cnt, 0, ctx->n_pad * ctx->n_pad * sizeof(uint32_t), ctx->stream);
  if (e != cudaSuccess) return e;
  if (n_entries == 0) return cudaSuccess;
  CountArgs a;
  a.mask = ctx->mask;
  a.ent_word = ctx->d_ent_word;
  a.ent_mask = ctx->d_ent_mask;
  a.tiles = ctx->d_tiles;
  a.cnt = ctx->d_cnt;
  a.NW = ctx->NW;
  a.n_pad = ctx->n_pad;
  a.n_entries = n_entries;
  uint64_t want = (uint64_t) (4 * ctx->n_sm + ctx->n_tiles - 1) / ctx->n_tiles;
  uint64_t maxs = (n_entries + kEB - 1) / kEB;
  a.n_splits = (uint32_t) (want < 1 ? 1 : (want > maxs ? maxs : want));
  if (a.n_splits > 65535) a.n_splits = 65535;
  k_mask_count<<<dim3(ctx->n_tiles, a.n_splits), 256, 0, ctx->stream>>>(a);
  return cudaGetLastError();
}
