// K2c dist_imma: the pair loop of gen_dist (ngsDist.cpp:333-364) for CALLED genotypes as an exact integer contraction on
// the int8 tensor cores (mma.sync.m16n8k32.s8 -> IMMA.16832.S8.S8).
//
// When every individual-site is either a hard call or missing -- genotype input (read_data.cpp:88-95) or --call_geno with
// N_thresh == call_thresh (gen_func.cpp:903-913; the default 0/0) -- its posterior triple is one of four states:
// one-hot(0), one-hot(1), one-hot(2) or the uniform triple of a missing entry (code 3).  The site term
//     sum_{g1,g2} score[g1][g2] p_i[g1] p_j[g2]            (ngsDist.cpp:351-353)
// then only depends on the two codes: a 4 x 4 table f(c_i, c_j).  With one-hot indicator bytes on the A side and the
// table column of the B side's code on the B side,
//     A_i[s][k] = w_s [c_i(s) == k]        B_j[s][k] = S * f(k, c_j(s))          k = 0..3, 4 bytes per site
//     acc(i,j) = sum_s sum_k A_i[s][k] B_j[s][k] = S * sum_s w_s f(c_i(s), c_j(s))
// is an int8 GEMM with K = 4 * n_sites and exact int32 accumulation; S = 2 (with --pairwise_del, where missing rows /
// columns of f are zero) or 18 (otherwise: the uniform 1/3 makes f a multiple of 1/18) turns f into small integers.
// w_s is the bootstrap multiplicity of the site (ngsDist.cpp:416-437), 1 for replicate 0, 0 for padding.
//
// HBM holds 2 bits per individual-site (0.25 B instead of 48 B of FP64 operands): codes[row block][64-site word][4][128]
// uint32, 16 sites per word.  The operands never exist in memory: each consumer warp expands its own mma fragments in
// registers -- A word = w << (8 * code) (3 integer ops), B word = one LDS from a 4-entry table -- so shared memory only
// carries the packed codes (4 KiB per 64 sites and tile) and the int8 tensor pipe, not the LSU, is the limit.
//
// Structure (one persistent CTA per SM, same scheduler as dist_dmma.cu): warp 8 = producer (dynamic (K-split, tile)
// units, cp.async.bulk + mbarrier ring of 64-site stages); warps 0..7 = consumers, 2 (M) x 4 (N), warp tile 64 x 32
// = 4 x 4 IMMA accumulators (64 registers).  Partials are int32 in fragment order; k_epilogue_int sums the splits in
// int64 and applies num = acc / S and the tail of gen_dist.  Everything is integer until that division: num, cnt and
// the model-0 distances are bit-exact whenever the reference's own sum is exact (no missing data, or --pairwise_del).
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "ngsd_internal.h"

namespace {

#ifndef NGSD_IMMA_CTAS
#define NGSD_IMMA_CTAS 1
#endif
constexpr int kStages = 8;
constexpr int kConsumerWarps = 8;
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kCodeBytes = 4 * 128 * 4;                       // one operand of one stage: [4 words][128 rows] uint32
constexpr int kMaskBytes = 128 * 8;                           // presence bits of one operand of one stage (COUNT only)
constexpr int kStageBytesMax = 2 * kCodeBytes + 64 + 2 * kMaskBytes;   // A codes + B codes + 64 site weights (uint8) [+ masks]
constexpr size_t kSmemBytes = (size_t) kStages * kStageBytesMax + 2 * kStages * sizeof(uint64_t) + kStages * 2 * sizeof(uint32_t) + 64 + 2 * 256 * 8 + 64;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// producer-side wait: back off between polls, the consumers on the same SM sub-partition need the issue slots
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(200);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void imma16832(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct ImmaArgs {
  const uint32_t *codes;        // [RB][NW][4][128]
  const uint64_t *mask;         // [RB][NW][128] presence bits (COUNT only)
  const uint8_t *wsite;         // [n_layers][NW * 64] per-site weights (0 for padding sites)
  const uint32_t *word_ids;     // [n_words] active 64-site words
  const uint32_t *word_layer;   // [n_words] weight layer of each entry
  const ngsd_tile *tiles;
  const uint32_t *split_begin;  // [n_splits + 1] boundaries in the word list
  uint32_t *sched;
  int32_t *partials;            // [n_units][16384]
  uint64_t NW;
  uint32_t n_tiles, n_units;
  uint32_t lut[4];              // B words: byte k of lut[c] = S * f(k, c)
  uint32_t pstride;             // ints per unit in `partials` (16384, or 32768 with the count tile behind the sum tile)
};

enum : uint32_t { kFirst = 1u, kLast = 2u, kExit = 8u };

// COUNT = false: the contraction described above.  COUNT = true: the same machinery computes ONLY the shared-site
// counts of --pairwise_del, cnt(i,j) = sum_s w_s m_i(s) m_j(s) (ngsDist.cpp:335-338,362), as an int8 GEMM with ONE byte
// per site (K = 32 sites per IMMA, a quarter of the instructions of the contraction): A' bytes = w_s m_i(s), B' bytes =
// m_j(s), expanded from the presence bit masks a byte (8 sites) at a time through two 256-entry tables.  On this path
// it replaces the AND+POPC kernel (mask_count.cu), which was the longer of the two when they ran side by side.  (One
// fused kernel with both accumulator sets needs > 168 registers -- the cap for a 9-warp CTA -- and spills.)
template <bool COUNT>
__global__ void __launch_bounds__(kThreads, NGSD_IMMA_CTAS) k_dist_imma(ImmaArgs a) {
  constexpr int kStageBytes = COUNT ? 2 * kMaskBytes + 64 : 2 * kCodeBytes + 64;
  constexpr int kOffW = COUNT ? 2 * kMaskBytes : 2 * kCodeBytes;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t) kStages * kStageBytes);
  uint64_t *empty = full + kStages;
  uint32_t *meta = reinterpret_cast<uint32_t *>(empty + kStages);   // [kStages][2] = {unit, flags}
  uint32_t *lut = meta + 2 * kStages;                               // [4]
  uint2 *lutA = reinterpret_cast<uint2 *>(lut + 4), *lutB = lutA + 256;   // [256] 8 presence bits -> 8 bytes of 0xFF / 0x01
  auto stageA = [&](int s) { return reinterpret_cast<const uint32_t *>(smem + (size_t) s * kStageBytes); };
  auto stageB = [&](int s) { return reinterpret_cast<const uint32_t *>(smem + (size_t) s * kStageBytes + kCodeBytes); };
  auto stageW = [&](int s) { return reinterpret_cast<const uint8_t *>(smem + (size_t) s * kStageBytes + kOffW); };
  auto stageMA = [&](int s) { return reinterpret_cast<const uint64_t *>(smem + (size_t) s * kStageBytes); };
  auto stageMB = [&](int s) { return reinterpret_cast<const uint64_t *>(smem + (size_t) s * kStageBytes + kMaskBytes); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x == 0) { lut[0] = a.lut[0]; lut[1] = a.lut[1]; lut[2] = a.lut[2]; lut[3] = a.lut[3]; }
  if (COUNT && threadIdx.x < 256) {
    const uint32_t n = threadIdx.x, lo = n & 15u, hi = n >> 4;
    const uint32_t one_lo = (lo & 1u) | ((lo & 2u) << 7) | ((lo & 4u) << 14) | ((lo & 8u) << 21);
    const uint32_t one_hi = (hi & 1u) | ((hi & 2u) << 7) | ((hi & 4u) << 14) | ((hi & 8u) << 21);
    lutB[n] = make_uint2(one_lo, one_hi);
    lutA[n] = make_uint2(one_lo * 0xFFu, one_hi * 0xFFu);
  }
  __syncthreads();

  int stage = 0;
  uint32_t phase = 0;

  if (warp == kConsumerWarps) {
    // ===== producer =====
    if (lane == 0) {
      for (;;) {
        const uint32_t u = atomicAdd(a.sched, 1u);
        if (u >= a.n_units) break;
        const uint32_t q = u / a.n_tiles, t = u - q * a.n_tiles;
        const ngsd_tile tl = a.tiles[t];
        const uint32_t c0 = a.split_begin[q], c1 = a.split_begin[q + 1];
        const uint32_t *Ab = a.codes + (uint64_t) tl.ti * a.NW * 512;
        const uint32_t *Bb = a.codes + (uint64_t) tl.tj * a.NW * 512;
        for (uint32_t c = c0; c < c1; c++) {
          const uint64_t word = a.word_ids[c];
          const uint8_t *wsrc = a.wsite + ((uint64_t) (a.word_layer[c] & 0x7FFFFFFFu) * a.NW + word) * 64;
          mbar_wait_sleep(&empty[stage], phase ^ 1);
          meta[stage * 2] = u;
          meta[stage * 2 + 1] = (c == c0 ? kFirst : 0u) | (c + 1 == c1 ? kLast : 0u);
          mbar_expect_tx(&full[stage], kStageBytes);
          unsigned char *dst = smem + (size_t) stage * kStageBytes;
          if (COUNT) {
            bulk_g2s(dst, a.mask + ((uint64_t) tl.ti * a.NW + word) * 128, kMaskBytes, &full[stage]);
            bulk_g2s(dst + kMaskBytes, a.mask + ((uint64_t) tl.tj * a.NW + word) * 128, kMaskBytes, &full[stage]);
          } else {
            bulk_g2s(dst, Ab + word * 512, kCodeBytes, &full[stage]);
            bulk_g2s(dst + kCodeBytes, Bb + word * 512, kCodeBytes, &full[stage]);
          }
          bulk_g2s(dst + kOffW, wsrc, 64, &full[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      mbar_wait(&empty[stage], phase ^ 1);
      meta[stage * 2 + 1] = kExit;
      mbar_arrive(&full[stage]);
    }
    return;
  }

  // ===== consumers: 2 x 4 warps, warp tile 64 (rows) x 32 (columns) =====
  const int wm = warp >> 2, wn = warp & 3, g = lane >> 2, q = lane & 3;
  const int rotA = (2 * q - 3) & 31, rotB = (2 * q - 2) & 31;       // bring this lane's code field to bits 3..4 / 2..3
  const int rotC = (8 * q - 3) & 31;                                // presence byte of this lane's 8 sites -> bits 3..10 (x 8 bytes)
  int acc[4][4][4];
  for (;;) {
    mbar_wait(&full[stage], phase);
    const uint32_t u = meta[stage * 2], fl = meta[stage * 2 + 1];
    if (fl & kExit) break;
    if (fl & kFirst) {
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++)
#pragma unroll
          for (int k = 0; k < 4; k++) acc[mi][ni][k] = 0;
    }
    if (COUNT) {
      // The K order of the count GEMM is free as long as both operands use it: lane q takes the 8 consecutive sites
      // 32 j + 8 q .. + 7 (K 4q..4q+3 = the first four, K 16+4q..+3 = the last four), i.e. ONE byte of a presence word,
      // expanded to its two fragment registers by one 8-byte table read; the 8 site weights are one 8-byte read too.
      const uint32_t *Ma = reinterpret_cast<const uint32_t *>(stageMA(stage)) + (wm * 64 + g) * 2;
      const uint32_t *Mb = reinterpret_cast<const uint32_t *>(stageMB(stage)) + (wn * 32 + g) * 2;
      const uint2 *W64 = reinterpret_cast<const uint2 *>(stageW(stage)) + q;
#pragma unroll
      for (int j = 0; j < 2; j++) {                                 // one IMMA k-step = 32 sites, one byte per site
        const uint2 ww = W64[j * 4];
        uint2 bf[4];
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {
          const uint32_t h = Mb[ni * 16 + j];
          bf[ni] = *reinterpret_cast<const uint2 *>(reinterpret_cast<const unsigned char *>(lutB) + (__funnelshift_r(h, h, rotC) & 0x7F8u));
        }
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {
          const uint32_t h0 = Ma[mi * 32 + j], h1 = Ma[mi * 32 + 16 + j];
          const uint2 e0 = *reinterpret_cast<const uint2 *>(reinterpret_cast<const unsigned char *>(lutA) + (__funnelshift_r(h0, h0, rotC) & 0x7F8u));
          const uint2 e1 = *reinterpret_cast<const uint2 *>(reinterpret_cast<const unsigned char *>(lutA) + (__funnelshift_r(h1, h1, rotC) & 0x7F8u));
          const uint32_t af[4] = {e0.x & ww.x, e1.x & ww.x, e0.y & ww.y, e1.y & ww.y};
#pragma unroll
          for (int ni = 0; ni < 4; ni++) imma16832(acc[mi][ni], af, bf[ni].x, bf[ni].y);
        }
      }
    }
    const uint32_t *As = stageA(stage) + wm * 64 + g, *Bs = stageB(stage) + wn * 32 + g;
    const uint8_t *Ws = stageW(stage) + q;
#pragma unroll
    for (int k = 0; k < (COUNT ? 0 : 4); k++) {                     // 16 sites per packed word
      // After the rotation the four code fields this lane needs from a word (sites q, q+4, q+8, q+12 of the 16) sit at
      // bits 3..4 (A: 8 * code) or 2..3 (B: 4 * code) of its four bytes; one mask per word isolates all of them.
      uint32_t ar[4][2], br[4];
#pragma unroll
      for (int mi = 0; mi < 4; mi++) {
        const uint32_t x0 = As[k * 128 + mi * 16], x1 = As[k * 128 + mi * 16 + 8];
        ar[mi][0] = __funnelshift_r(x0, x0, rotA) & 0x18181818u;
        ar[mi][1] = __funnelshift_r(x1, x1, rotA) & 0x18181818u;
      }
#pragma unroll
      for (int ni = 0; ni < 4; ni++) {
        const uint32_t y = Bs[k * 128 + ni * 8];
        br[ni] = __funnelshift_r(y, y, rotB) & 0x0C0C0C0Cu;
      }
#pragma unroll
      for (int p = 0; p < 2; p++) {                                 // one IMMA k-step = 8 sites = K 32
        const uint32_t w0 = Ws[k * 16 + p * 8], w1 = Ws[k * 16 + p * 8 + 4];
        uint32_t bf[4][2];
#pragma unroll
        for (int ni = 0; ni < 4; ni++) {                            // B word = table column of the code: one LDS
          bf[ni][0] = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const unsigned char *>(lut) + __byte_perm(br[ni], 0, 0x4440 + 2 * p));
          bf[ni][1] = *reinterpret_cast<const uint32_t *>(reinterpret_cast<const unsigned char *>(lut) + __byte_perm(br[ni], 0, 0x4441 + 2 * p));
        }
#pragma unroll
        for (int mi = 0; mi < 4; mi++) {                            // A word = weight << (8 * code); shf.wrap reads 5 bits only
          uint32_t af[4];
          af[0] = __funnelshift_l(0u, w0, ar[mi][0] >> (16 * p));
          af[1] = __funnelshift_l(0u, w0, ar[mi][1] >> (16 * p));
          af[2] = __funnelshift_l(0u, w1, ar[mi][0] >> (16 * p + 8));
          af[3] = __funnelshift_l(0u, w1, ar[mi][1] >> (16 * p + 8));
#pragma unroll
          for (int ni = 0; ni < 4; ni++) imma16832(acc[mi][ni], af, bf[ni][0], bf[ni][1]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    if (++stage == kStages) { stage = 0; phase ^= 1; }
    if (fl & kLast) {
      int4 *dst = reinterpret_cast<int4 *>(a.partials + (uint64_t) u * a.pstride) + (warp * 16) * 32 + lane;
#pragma unroll
      for (int mi = 0; mi < 4; mi++)
#pragma unroll
        for (int ni = 0; ni < 4; ni++)      // the count launch writes the second tile of the unit's slot
          dst[(COUNT ? NGSD_TILE_ELEMS / 4 : 0) + (mi * 4 + ni) * 32] = make_int4(acc[mi][ni][0], acc[mi][ni][1], acc[mi][ni][2], acc[mi][ni][3]);
    }
  }
}

// register-only IMMA loop: the int8 tensor issue-rate ceiling of mma.sync on this part (roofline denominator of K2c)
__global__ void k_imma_peak(int *out, int iters) {
  int c[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0;
  const uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u};
  const uint32_t b0 = threadIdx.x + 1u, b1 = 5u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) imma16832(c[i], a, b0, b1);
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 0x12345678) out[0] = s;
}

struct EpiIntArgs {
  const int32_t *partials;    // [n_splits][n_tiles][pstride] fragment order (sum tile, then count tile when in_kernel_cnt)
  const double *split_w;      // [n_splits] integer weight of each split (bootstrap block cache) or nullptr
  uint32_t pstride;
  int in_kernel_cnt;
  int sum_row_major;          // sum tile written by dist_umma.cu: plain [128][128] instead of mma.sync fragment order
  const ngsd_tile *tiles;
  const uint32_t *cnt;        // [n_pad][n_pad] (K3) or nullptr
  double *out, *num;
  uint64_t *cntout;
  uint64_t n_ind, n_pad, const_cnt, tot_sites;
  uint32_t n_splits, n_tiles;
  int evol_model;
  double scale;               // S
};

// grid (n_tiles, 16); block 256: one thread per int4 slot of the tile; splits summed in int64 (order-free, exact)
__global__ void __launch_bounds__(256) k_epilogue_int(EpiIntArgs a) {
  const uint32_t t = blockIdx.x;
  const ngsd_tile tl = a.tiles[t];
  const uint32_t e = blockIdx.y * 256 + threadIdx.x;     // int4 index inside the tile: (warp*16 + frag)*32 + lane
  const int lane = e & 31, frag = (e >> 5) & 15, warp = e >> 9;
  const int wm = warp >> 2, wn = warp & 3, mi = frag >> 2, ni = frag & 3;
  const int row0 = wm * 64 + mi * 16 + (lane >> 2), col0 = wn * 32 + ni * 8 + 2 * (lane & 3);
  const int4 *src = reinterpret_cast<const int4 *>(a.partials + (uint64_t) t * a.pstride) + e;
  const uint64_t stride = (uint64_t) a.n_tiles * (a.pstride / 4);
  long long s[4] = {0, 0, 0, 0}, c[4] = {0, 0, 0, 0};
  for (uint32_t q = 0; q < a.n_splits; q++) {
    long long m = 1;
    if (a.split_w) {                     // per-block partials x block multiplicities: integers, exact
      m = (long long) a.split_w[q];
      if (m == 0) continue;
    }
    int4 v;
    if (a.sum_row_major) {
      const int32_t *base = a.partials + ((uint64_t) q * a.n_tiles + t) * a.pstride;
      const int2 lo = *reinterpret_cast<const int2 *>(base + row0 * NGSD_TILE + col0), hi = *reinterpret_cast<const int2 *>(base + (row0 + 8) * NGSD_TILE + col0);
      v = make_int4(lo.x, lo.y, hi.x, hi.y);
    } else {
      v = src[(uint64_t) q * stride];
    }
    s[0] += m * v.x; s[1] += m * v.y; s[2] += m * v.z; s[3] += m * v.w;
    if (a.in_kernel_cnt) {
      int4 w;
      if (a.sum_row_major) {
        const int32_t *base = a.partials + ((uint64_t) q * a.n_tiles + t) * a.pstride + NGSD_TILE_ELEMS;
        const int2 lo = *reinterpret_cast<const int2 *>(base + row0 * NGSD_TILE + col0), hi = *reinterpret_cast<const int2 *>(base + (row0 + 8) * NGSD_TILE + col0);
        w = make_int4(lo.x, lo.y, hi.x, hi.y);
      } else {
        w = src[(uint64_t) q * stride + NGSD_TILE_ELEMS / 4];
      }
      c[0] += m * w.x; c[1] += m * w.y; c[2] += m * w.z; c[3] += m * w.w;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const uint64_t i = (uint64_t) tl.ti * NGSD_TILE + row0 + (k >> 1) * 8, j = (uint64_t) tl.tj * NGSD_TILE + col0 + (k & 1);
    if (i >= j || j >= a.n_ind) continue;
    const double num = (double) s[k] / a.scale;
    uint64_t cnt = a.in_kernel_cnt ? (uint64_t) c[k] : (a.cnt ? (uint64_t) a.cnt[i * a.n_pad + j] : a.const_cnt);
    if (a.num) a.num[i * a.n_ind + j] = a.num[j * a.n_ind + i] = num;
    if (a.cntout) a.cntout[i * a.n_ind + j] = a.cntout[j * a.n_ind + i] = cnt;
    if (a.tot_sites > 0) cnt = a.tot_sites;
    double d = num / (double) cnt;
    if (a.evol_model == 1) d = -log(1 - d);
    else if (a.evol_model == 2) d = -log(1 - (d * 4 / 3)) * 3 / 4;
    a.out[i * a.n_ind + j] = a.out[j * a.n_ind + i] = d;
  }
}

__global__ void k_zero_diag_int(double *out, double *num, uint64_t *cnt, uint64_t n) {
  const uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i * n + i] = 0.0;
  if (num) num[i * n + i] = 0.0;
  if (cnt) cnt[i * n + i] = 0;
}

}  // namespace

// f(k, c) * S as int8 bytes; returns false when the score matrix does not give small integers (-> FP64 path)
bool ngsd_int_lut(const double *score, bool pairwise_del, uint32_t lut[4], double *scale, int *max_byte) {
  const double S = pairwise_del ? 2.0 : 18.0;
  double f[4][4];
  double rs[3] = {0, 0, 0}, cs[3] = {0, 0, 0}, tot = 0;
  for (int k = 0; k < 3; k++)
    for (int c = 0; c < 3; c++) { f[k][c] = score[k * 3 + c]; rs[k] += score[k * 3 + c]; cs[c] += score[k * 3 + c]; tot += score[k * 3 + c]; }
  for (int k = 0; k < 3; k++) { f[k][3] = pairwise_del ? 0.0 : rs[k] / 3.0; f[3][k] = pairwise_del ? 0.0 : cs[k] / 3.0; }
  f[3][3] = pairwise_del ? 0.0 : tot / 9.0;
  int mx = 0;
  for (int c = 0; c < 4; c++) {
    uint32_t wv = 0;
    for (int k = 0; k < 4; k++) {
      const double v = f[k][c] * S, r = nearbyint(v);
      if (fabs(v - r) > 1e-9 || r < 0 || r > 127) return false;
      if (!pairwise_del && k < 3 && ((int) r) % 3) return false;   // dist_umma.cu's three-plane form carries S / 3 on the B side
      wv |= (uint32_t) (int) r << (8 * k);
      if ((int) r > mx) mx = (int) r;
    }
    lut[c] = wv;
  }
  *scale = S;
  *max_byte = mx;
  return true;
}

int ngsd_imma_ctas_per_sm() { return NGSD_IMMA_CTAS; }

// NGSD_IMMA_SYNC=1 keeps the mma.sync contraction (A/B comparison against the tcgen05 path of dist_umma.cu)
bool ngsd_use_umma() {
  static const bool off = getenv("NGSD_IMMA_SYNC") != nullptr;
  return !off;
}

cudaError_t ngsd_launch_dist_imma(ngsd_ctx *ctx, uint32_t n_units, int grid, bool count) {
  static bool attr_set[64] = {};
  if (!attr_set[ctx->device & 63]) {
    const void *fns[2] = {(const void *) k_dist_imma<false>, (const void *) k_dist_imma<true>};
    for (const void *f : fns) {
      cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBytes);
      if (e != cudaSuccess) return e;
    }
    attr_set[ctx->device & 63] = true;
  }
  ImmaArgs a;
  a.codes = ctx->codes;
  a.mask = ctx->mask;
  a.pstride = count ? 2 * NGSD_TILE_ELEMS : NGSD_TILE_ELEMS;
  a.wsite = ctx->d_wsite;
  a.word_ids = ctx->d_word_ids;
  a.word_layer = ctx->d_word_layer;
  a.tiles = ctx->d_tiles;
  a.split_begin = ctx->d_split_begin;
  a.sched = ctx->d_sched;
  a.partials = reinterpret_cast<int32_t *>(ctx->cur_partials);
  a.NW = ctx->NW;
  a.n_tiles = ctx->n_tiles;
  a.n_units = n_units;
  for (int k = 0; k < 4; k++) a.lut[k] = ctx->int_lut[k];
  cudaError_t e = cudaMemsetAsync(ctx->d_sched, 0, sizeof(uint32_t), ctx->stream);
  if (e != cudaSuccess) return e;
  if (ngsd_use_umma()) {
    e = ngsd_launch_dist_umma(ctx, n_units, std::min(grid, ctx->n_sm), a.pstride, false);
    if (e == cudaSuccess && count) {
      e = cudaMemsetAsync(ctx->d_sched, 0, sizeof(uint32_t), ctx->stream);
      if (e == cudaSuccess) e = ngsd_launch_dist_umma(ctx, n_units, std::min(grid, ctx->n_sm), a.pstride, true);
    }
    return e;
  }
  k_dist_imma<false><<<grid, kThreads, kSmemBytes, ctx->stream>>>(a);
  if (count) {      // second pass of the same unit list: the shared-site counts
    e = cudaMemsetAsync(ctx->d_sched, 0, sizeof(uint32_t), ctx->stream);
    if (e != cudaSuccess) return e;
    k_dist_imma<true><<<grid, kThreads, kSmemBytes, ctx->stream>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t ngsd_launch_epilogue_int(ngsd_ctx *ctx, uint32_t n_splits, uint64_t const_cnt, bool use_cnt, bool in_kernel_cnt) {
  EpiIntArgs a;
  a.partials = reinterpret_cast<const int32_t *>(ctx->cur_partials);
  a.split_w = ctx->cur_split_w;
  a.pstride = in_kernel_cnt ? 2 * NGSD_TILE_ELEMS : NGSD_TILE_ELEMS;
  a.in_kernel_cnt = in_kernel_cnt ? 1 : 0;
  a.sum_row_major = ngsd_use_umma() ? 1 : 0;
  a.tiles = ctx->d_tiles;
  a.cnt = (use_cnt && !in_kernel_cnt) ? ctx->d_cnt : nullptr;
  a.out = ctx->d_out;
  a.num = ctx->d_num;
  a.cntout = ctx->d_cntout;
  a.n_ind = ctx->n_ind;
  a.n_pad = ctx->n_pad;
  a.const_cnt = const_cnt;
  a.tot_sites = ctx->cfg.tot_sites;
  a.n_splits = n_splits;
  a.n_tiles = ctx->n_tiles;
  a.evol_model = ctx->cfg.evol_model;
  a.scale = ctx->int_scale;
  if (ctx->shard_world > 1) {   // entries of tiles owned by other ranks must read as 0 (ngsd_set_tile_shard)
    const uint64_t n2 = ctx->n_ind * ctx->n_ind;
    cudaError_t e = cudaMemsetAsync(ctx->d_out, 0, n2 * sizeof(double), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_num, 0, n2 * sizeof(double), ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_cntout, 0, n2 * sizeof(uint64_t), ctx->stream);
    if (e != cudaSuccess) return e;
  }
  k_zero_diag_int<<<(unsigned) ((ctx->n_ind + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_out, ctx->d_num, ctx->d_cntout, ctx->n_ind);
  k_epilogue_int<<<dim3(ctx->n_tiles, 16), 256, 0, ctx->stream>>>(a);
  return cudaGetLastError();
}

extern "C" int ngsd_probe_int8_tmacs(int device, double *imma_tmacs) {
  if (cudaSetDevice(device) != cudaSuccess) return NGSD_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return NGSD_ERR_CUDA;
  int *d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return NGSD_ERR_CUDA;
  const int iters = 8192, warps = 16, nsm = prop.multiProcessorCount;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < 4; r++) {
    cudaEventRecord(e0);
    k_imma_peak<<<nsm, warps * 32>>>(d, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return NGSD_ERR_CUDA; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *imma_tmacs = (double) nsm * warps * iters * 8 * 4096.0 / (best * 1e-3) * 1e-12;
  return NGSD_OK;
}
