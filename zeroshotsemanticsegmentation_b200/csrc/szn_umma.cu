// tcgen05 implicit-GEMM convolution kernels for sm_100a (forward, data-gradient, weight-gradient).
//
// One persistent, warp-specialised kernel template covers the GEMM shapes of the VGG16-FCN32s trunk
// (reference: models.py:43-98 forward, autograd backward of the same layers):
//
//   MODE 0  forward   D[pixel, co] = sum_{tap, ci} X[pixel + tap - pad, ci] * Wt[co, tap, ci]
//           A = activations, K-major (NHWC: channels contiguous), loaded as 4-D TMA boxes
//               (KC channels x TW x TH pixels) at the tap-shifted coordinate; out-of-image pixels are
//               zero-filled by TMA, which implements the conv padding.
//           B = weights [Cout][taps*Cin], K-major, 2-D TMA boxes.
//           dgrad runs on the same path: it is the forward conv of dY with the transposed, flipped weights
//           (szn_pack_weight_dgrad) and padding R-1-pad; its epilogue fuses the ReLU gate, the Dropout2d scale
//           and the bias gradient (column sums) of the layer that produced the conv's input.
//   MODE 2  wgrad     D[co, (tap, ci)] = sum_{pixel} dY[pixel, co] * X[pixel + tap - pad, ci]
//           both operands MN-major (the reduction runs over pixels), fetched with 5-D TMA boxes (one box lands all
//           128-byte channel groups of an operand), split-K over pixel chunks, reduced into dW[Cout][taps*Cin] by
//           TMA fp32 reduce-add; two 128-row output sub-tiles share every B stage when Cout >= 256.
//
// Tiles: UMMA M = 128 (rows = pixels of a TW x TH patch, or output channels for wgrad), N = block_n
// (32..256), K per stage = 128 bytes of the contraction dim (64 bf16 / 32 tf32), SWIZZLE_128B smem
// layouts (SWIZZLE_128B_BASE32B for tf32 MN-major) written by TMA and consumed through shared-memory
// descriptors, fp32 accumulators in TMEM (two buffers: the epilogue of a tile overlaps the next main loop).
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue
// (TMEM -> registers -> swizzled staging rows -> TMA store / reduce-add).  192 threads, one CTA per SM.
//
// SPLIT (dtype SZN_F32X3, "fp32-grade"): every operand arrives as two bf16 planes hi / lo (szn_store.cuh); a stage holds
// A_hi | A_lo | B_hi | B_lo and every K step issues three kind::f16 MMAs hi*hi + lo*hi + hi*lo into the same fp32
// accumulator (the lo*lo term is below 2^-18 of the product).  The epilogue splits its fp32 results into the two output
// planes again.  Same bytes per element as the tf32 path, 1.5x its MMA time, ~2^-16 instead of 2^-11 per product.
#include "szn_internal.h"
#include "szn_ptx.cuh"
#include <stdlib.h>
#include <string.h>

namespace szn {

struct UmmaParams {
  int tiles_x, tiles_y, B;  // A-side spatial tiling (M tiles for MODE 0/1, K chunks for MODE 2)
  int TW, TH;               // tile = TW x TH pixels
  int H, W;                 // MODE 0/1: extent of the output image
  int R, S, pad;
  int Ck;       // MODE 0/1: contraction channels per tap
  int kchunks;  // ceil(Ck / KC)
  int N;        // valid GEMM columns
  int M;        // MODE 2: valid GEMM rows (Cout)
  int Cin;      // MODE 2: channels per tap inside the N index
  int block_n, n_tiles, m_tiles, stages, splits, tmem_cols;
  int total_tiles;  // n_tiles * m_tiles (* splits): work items of the persistent tile loop
  int m_fast;  // MODE 0: tile order (0: N tiles fastest, share A through L2; 1: M tiles fastest, share the weight slice)
  int mpair;  // MODE 2: 128-row output sub-tiles per work item (2: two accumulators share every B stage)
  int nbuf;   // accumulator buffers in TMEM (2 unless one tile needs all 512 columns)
  int acc_cols;  // TMEM column stride between the accumulators of a tile (MODE 2 tile pairs, MODE 0 nacc = 2)
  int nacc;      // MODE 0: accumulators per tile (see launch(): shorter fp32 accumulation chains), summed by the epilogue
  int gpt, b_boxes, ksteps;  // MODE 2: 128-byte column groups per B box, B boxes per stage, MMAs (K steps) per stage
  long long ldo;          // channels per output row (= its row stride in elements; split: stride of one plane pair / 2)
  long long ld_mask;      // row stride of mask_ref in elements of T (split: 2 * channels, the hi plane comes first)
  int a_lo_goff, b_lo_goff;  // SPLIT, MODE 2: channel-group offset of the lo plane inside a pixel row of dy / x
  int a_lo_ch;               // SPLIT, MODE 0 im2col: channel offset of the lo plane inside a pixel row of the input (= ldx)
  const float* bias;      // [N] or null
  const float* scale;     // [B][scale_ld] per-(image, channel) multiplier (Dropout2d) or null
  int scale_ld;
  int pix_per_image;      // rows per image for the scale lookup (H*W, or the original H*W when flattened)
  const void* mask_ref;   // dgrad: activation whose sign gates the gradient (ReLU backward) or null
  int relu, out_fp32;
  int vec_ok;  // bias / scale rows are 16-byte aligned
  float* col_sum;  // dgrad: += column sums of the stored output (the producer layer's bias gradient) or null
  int cs_local;    // col_sum && one N tile: the CTA sums its tiles' columns in shared memory and adds them to col_sum once
  unsigned int* sched;  // {next tile, CTAs done}: global work counter of this launch (self-resetting, see launch())
  int im2col;           // MODE 0: M tiles are runs of 128 consecutive output pixels, A fetched by im2col-mode TMA
  long long m_total;    // MODE 0 im2col: output pixels B * Ho * Wo
  int static_tiles;     // tuning / A-B switch (SZN_STATIC_TILES=1): round-robin tile list instead of the work counter
};

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// out[j] = src[j] for j < nvalid (0 beyond); all loads are issued before any use
template <int NV>
__device__ __forceinline__ void load_row(const float* __restrict__ src, int cw, int nvalid, bool vec, float (&out)[NV]) {
  if (vec) {
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      if (4 * j < cw) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + j);
        out[4 * j] = v.x, out[4 * j + 1] = v.y, out[4 * j + 2] = v.z, out[4 * j + 3] = v.w;
      } else {
        out[4 * j] = out[4 * j + 1] = out[4 * j + 2] = out[4 * j + 3] = 0.f;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < NV; ++j) out[j] = (j < nvalid) ? __ldg(src + j) : 0.f;
  }
}

struct TileCoord {
  int n0, b, x0, y0, m0, q_begin, n_iters;
};

// work item -> tile coordinates; n fastest so that CTAs running side by side share the A (pixel) tile through L2.
// One set of divisions per TILE (not per pipeline stage).
template <int MODE>
__device__ __forceinline__ TileCoord decode_tile(const UmmaParams& p, int tile) {
  TileCoord t;
  int n_tile, idx;
  if (MODE == 0 && p.m_fast) {  // pixel tiles fastest: CTAs running side by side share the B (weight) slice instead
    const int mt = p.total_tiles / p.n_tiles;
    n_tile = tile / mt, idx = tile - n_tile * mt;
  } else {
    n_tile = tile % p.n_tiles, idx = tile / p.n_tiles;
  }
  t.n0 = n_tile * p.block_n;
  t.b = t.x0 = t.y0 = t.m0 = t.q_begin = 0;
  const int tiles_per_img = p.tiles_x * p.tiles_y;
  if (MODE == 2) {
    const int m_tile = idx % p.m_tiles, split = idx / p.m_tiles;
    t.m0 = m_tile * 128 * p.mpair;
    const int total_q = tiles_per_img * p.B;
    const int per = (total_q + p.splits - 1) / p.splits;
    t.q_begin = split * per;
    int q_end = t.q_begin + per;
    if (q_end > total_q) q_end = total_q;
    t.n_iters = q_end - t.q_begin;
    if (t.n_iters < 0) t.n_iters = 0;
  } else if (p.im2col) {
    // a run of 128 output pixels starting at linear index m0 = idx * 128 (kept in q_begin); (x0, y0, b) = its first pixel
    const long long m0 = (long long)idx * 128;
    const int hw = p.H * p.W;
    t.b = (int)(m0 / hw);
    const int r = (int)(m0 - (long long)t.b * hw);
    t.y0 = r / p.W;
    t.x0 = r - t.y0 * p.W;
    t.q_begin = idx;
    t.n_iters = p.R * p.S * p.kchunks;
  } else {
    t.b = idx / tiles_per_img;
    const int r = idx - t.b * tiles_per_img;
    const int ty = r / p.tiles_x;
    t.y0 = ty * p.TH;
    t.x0 = (r - ty * p.tiles_x) * p.TW;
    t.n_iters = p.R * p.S * p.kchunks;
  }
  return t;
}

// Persistent, warp-specialised implicit-GEMM kernel.  grid = min(#tiles, #SMs).  Tiles are handed out DYNAMICALLY: a CTA's
// first tile is blockIdx.x, every further one comes from a global atomic counter, fetched by the producer warp (one tile
// ahead, so the atomic's latency hides behind the loads) and passed to the MMA and epilogue warps through a 4-deep
// shared-memory queue.  A CTA that starts late or runs slowly -- its SM is shared with NCCL's all-reduce kernels, or was
// still busy with the previous kernel's tail -- simply takes fewer tiles instead of setting the kernel's duration, which is
// what a static round-robin list did (8-GPU step 22.5 -> 24.5 ms in round 1).  Pipelines:
//   tile queue  sqf[q]/sqe[q]          TMA producer  -> MMA issuer, epilogue warps
//   smem ring   full[s]/empty[s]       TMA producer  -> MMA issuer
//   TMEM        accf[2]/acce[2]        MMA issuer    -> epilogue (two accumulator buffers: the epilogue of tile i
//                                                       overlaps the main loop of tile i+1)
//   staging     cp.async.bulk groups   epilogue      -> TMA store (two 16 KB swizzled buffers; coalesced, clipped by
//                                                       the tensor map, fp32 add-reduction for the split-K wgrad)
template <typename T, int MODE, bool SPLIT>
__global__ void __launch_bounds__(192, 1)
umma_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const UmmaParams p) {
  static_assert(!SPLIT || sizeof(T) == 2, "the split format is made of bf16 planes");
  constexpr bool TF32 = sizeof(T) == 4;
  constexpr int NPL = SPLIT ? 2 : 1;        // operand planes per stage
  constexpr int KC = 128 / (int)sizeof(T);  // contraction elements per stage
  constexpr int UK = KC / 4;                // UMMA K (16 bf16 / 8 tf32): 4 MMAs per stage
  constexpr int A_BYTES = 128 * 128;
  constexpr int STAGING_BYTES = 128 * 128;
  constexpr bool A_MN = (MODE == 2);
  constexpr bool B_MN = (MODE == 2);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int block_n = p.block_n;
  const int mp = (MODE == 2) ? p.mpair : 1;
  const int a_bytes = A_BYTES * mp;          // one plane of A
  const int b_bytes = block_n * 128;          // one plane of B
  const int stage_bytes = NPL * (a_bytes + b_bytes);  // A_hi | A_lo | B_hi | B_lo
  const int stages = p.stages;
  uint8_t* staging = smem + stages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + 2 * STAGING_BYTES);
  uint64_t* empty = full + 8;
  uint64_t* accf = empty + 8;  // [2] accumulator buffer complete
  uint64_t* acce = accf + 2;   // [2] accumulator buffer drained
  constexpr int SQ = 4;        // depth of the tile queue
  uint64_t* sqf = acce + 2;    // [SQ] tile index published by the producer warp
  uint64_t* sqe = sqf + SQ;    // [SQ] tile index read by the MMA warp and the four epilogue warps
  int* sq_tile = reinterpret_cast<int*>(sqe + SQ);  // [SQ] tile index, -1 = no more work
  uint32_t* tptr = reinterpret_cast<uint32_t*>(sq_tile + SQ);
  // [256] per-CTA column sums of the stored tiles (p.cs_local): conv1_2's data gradient has 31.5k tiles at B = 8, and one
  // RED per (warp, column, tile) meant 8 M atomics onto the two cache lines that hold 64 bias gradients
  float* cs_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full) + 384);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) cs_smem[i] = 0.f;

  const int warp = warp_idx(), lane = threadIdx.x & 31;  // provably warp-uniform (see elect_one)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&accf[i], 1);
      mbar_init(&acce[i], 4);  // one arrival per epilogue warp
    }
    for (int i = 0; i < SQ; ++i) {
      mbar_init(&sqf[i], 1);
      mbar_init(&sqe[i], 5);  // MMA warp + four epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tptr, (uint32_t)(p.nbuf * p.tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;

  const int rows_a = p.TW * p.TH;  // rows written by one pixel box
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    // The whole warp walks the loop (uniform control flow); one elected lane issues.  All index decompositions inside a
    // tile are carried incrementally (a runtime integer division costs ~100 cycles; 19 of them per wgrad stage made the
    // producer the bottleneck in the first profile).
    // MODE 2 operands are fetched with 5-D boxes {KC channels, TW, TH, 1, G channel groups}: one box lands G column
    // groups (each rows_a x 128 B, back to back) instead of one 4 KB box per group -- the TMA unit pays a fixed cost per
    // box, and twelve small boxes per stage made the wgrad load-bound.
    const uint32_t a_tx = (MODE == 2) ? (uint32_t)(mp * (128 / KC) * rows_a * 128) : (uint32_t)(rows_a * 128);
    const uint32_t box_tx = (uint32_t)(p.gpt * rows_a * 128);  // MODE 2: bytes of one B box
    int s = 0;
    uint32_t ph = 0;
    int tile = blockIdx.x, nxt = 0;
    for (uint32_t qi = 0;; ++qi) {
      // publish this tile (or the end marker) to the other roles
      const int slot = qi & (SQ - 1);
      mbar_wait(&sqe[slot], ((qi / SQ) & 1u) ^ 1u);
      const bool live = tile < p.total_tiles;
      if (lane == 0) {
        sq_tile[slot] = live ? tile : -1;
        mbar_arrive(&sqf[slot]);
        // claim the next tile right away: the atomic's round trip hides behind this tile's loads
        if (live) nxt = p.static_tiles ? tile + (int)gridDim.x : (int)atomicAdd(p.sched, 1u) + (int)gridDim.x;
      }
      __syncwarp();
      if (!live) break;
      const TileCoord t = decode_tile<MODE>(p, tile);
      int tap = 0, cc = 0, r = 0, sx = 0;  // MODE 0: filter tap (r, sx) and channel chunk
      int bb = 0, py0 = 0, px0 = 0;        // MODE 2: pixel chunk (image, tile origin)
      int g_cg[4], g_dx[4], g_dy[4];       // MODE 2: per B box: first channel group and tap shift
      int n_bbox = 0;                      // MODE 2: boxes of this tile that lie inside the filter (last N tile may be short)
      if (MODE == 2) {
        const int bq = t.q_begin / tiles_per_img;
        const int tt = t.q_begin - bq * tiles_per_img;
        const int ty = tt / p.tiles_x;
        bb = bq, py0 = ty * p.TH, px0 = (tt - ty * p.tiles_x) * p.TW;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int nn = t.n0 + j * p.gpt * KC;
          const int tp = nn / p.Cin;
          g_cg[j] = (nn - tp * p.Cin) / KC;
          const int rr = tp / p.S;
          g_dy[j] = rr - p.pad, g_dx[j] = tp - rr * p.S - p.pad;
          if (j < p.b_boxes && tp < p.R * p.S) n_bbox = j + 1;
        }
      }
      const uint32_t b_tx = (MODE == 2) ? (uint32_t)n_bbox * box_tx : (uint32_t)(block_n * 128);
      for (int it = 0; it < t.n_iters; ++it) {
        mbar_wait(&empty[s], ph ^ 1u);
        if (elect_one()) {
          uint8_t* a_dst = smem + s * stage_bytes;
          uint8_t* b_dst = a_dst + NPL * a_bytes;
          mbar_expect_tx(&full[s], NPL * (a_tx + b_tx));
          if (MODE == 0 && p.im2col) {
            // base pixel of the run relative to the tensor (-pad = first output pixel), tap (r, sx) as the im2col offset;
            // split: a pixel row is [hi | lo], the lo plane's channels simply follow at + Ck
            tma_load_im2col_4d(a_dst, &tmA, &full[s], cc * KC, t.x0 - p.pad, t.y0 - p.pad, t.b, (uint16_t)sx, (uint16_t)r);
            if (SPLIT) {
              tma_load_im2col_4d(a_dst + a_bytes, &tmA, &full[s], p.a_lo_ch + cc * KC, t.x0 - p.pad, t.y0 - p.pad, t.b, (uint16_t)sx,
                                 (uint16_t)r);
              tma_load_3d(b_dst, &tmB, &full[s], tap * p.Ck + cc * KC, t.n0, 0);
              tma_load_3d(b_dst + b_bytes, &tmB, &full[s], tap * p.Ck + cc * KC, t.n0, 1);
            } else {
              tma_load_2d(b_dst, &tmB, &full[s], tap * p.Ck + cc * KC, t.n0);
            }
          } else if (MODE == 0) {
            if (SPLIT) {  // planes are the 5th (A) / 3rd (B) tensor-map dimension
              tma_load_5d(a_dst, &tmA, &full[s], cc * KC, t.x0 + sx - p.pad, t.y0 + r - p.pad, t.b, 0);
              tma_load_5d(a_dst + a_bytes, &tmA, &full[s], cc * KC, t.x0 + sx - p.pad, t.y0 + r - p.pad, t.b, 1);
              tma_load_3d(b_dst, &tmB, &full[s], tap * p.Ck + cc * KC, t.n0, 0);
              tma_load_3d(b_dst + b_bytes, &tmB, &full[s], tap * p.Ck + cc * KC, t.n0, 1);
            } else {
              tma_load_4d(a_dst, &tmA, &full[s], cc * KC, t.x0 + sx - p.pad, t.y0 + r - p.pad, t.b);
              tma_load_2d(b_dst, &tmB, &full[s], tap * p.Ck + cc * KC, t.n0);
            }
          } else {
            tma_load_5d(a_dst, &tmA, &full[s], 0, px0, py0, bb, t.m0 / KC);
            if (SPLIT) tma_load_5d(a_dst + a_bytes, &tmA, &full[s], 0, px0, py0, bb, p.a_lo_goff + t.m0 / KC);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < n_bbox) {
                tma_load_5d(b_dst + j * box_tx, &tmB, &full[s], 0, px0 + g_dx[j], py0 + g_dy[j], bb, g_cg[j]);
                if (SPLIT)
                  tma_load_5d(b_dst + b_bytes + j * box_tx, &tmB, &full[s], 0, px0 + g_dx[j], py0 + g_dy[j], bb,
                              p.b_lo_goff + g_cg[j]);
              }
          }
        }
        __syncwarp();
        if (++s == stages) s = 0, ph ^= 1u;
        if (MODE == 2) {
          px0 += p.TW;
          if (px0 >= p.tiles_x * p.TW) {
            px0 = 0, py0 += p.TH;
            if (py0 >= p.tiles_y * p.TH) py0 = 0, ++bb;
          }
        } else {
          if (++cc == p.kchunks) {
            cc = 0, ++tap;
            if (++sx == p.S) sx = 0, ++r;
          }
        }
      }
      tile = __shfl_sync(0xffffffffu, nxt, 0);
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    // whole warp in the loop, one elected lane issues: back-to-back UTCHMMA without the ELECT/BRA.U.ANY loops (see elect_one)
    const uint32_t idesc = umma_idesc(TF32 ? 2 : 1, A_MN ? 1 : 0, B_MN ? 1 : 0, 128, block_n);
    int s = 0;
    uint32_t ph = 0;
    uint32_t local = 0;  // tiles this CTA has started: accumulator buffer = local & 1
    for (uint32_t qi = 0;; ++qi) {
      const int slot = qi & (SQ - 1);
      mbar_wait(&sqf[slot], (qi / SQ) & 1u);
      const int tile = sq_tile[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sqe[slot]);
      if (tile < 0) break;
      const TileCoord t = decode_tile<MODE>(p, tile);
      if (t.n_iters == 0) continue;
      const uint32_t buf = p.nbuf == 2 ? (local & 1u) : 0u, aph = p.nbuf == 2 ? ((local >> 1) & 1u) : (local & 1u);
      ++local;
      mbar_wait(&acce[buf], aph ^ 1u);  // the epilogue has drained this accumulator buffer
      tc_fence_after();
      const uint32_t dacc = tmem + buf * (uint32_t)p.tmem_cols;
      for (int it = 0; it < t.n_iters; ++it) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
          const uint32_t b_addr = a_addr + NPL * a_bytes;
          // K-major: 8-row groups 1024 B apart, K advances 32 B inside the 128 B swizzled row.
          // MN-major: 128 B-wide column groups LBO = rows_a*128 B apart (as the 5-D TMA box lays them down), K advances UK
          // rows of 128 B per MMA; the K rows come in groups SBO apart: 8 rows / 1024 B for 16-bit operands (SWIZZLE_128B),
          // 4 rows / 512 B for tf32 (SWIZZLE_128B_BASE32B, the only MN-major layout tcgen05 takes for 32-bit operands).
          constexpr uint32_t MN_LAYOUT = TF32 ? 1u : 2u, MN_SBO = TF32 ? 512u : 1024u;
          const uint32_t lbo = (uint32_t)rows_a * 128u;
          const int ksteps = (MODE == 2) ? p.ksteps : 4;
          if (MODE == 2 && mp == 2) {
            // two 128-row sub-tiles share the B stage (never with SPLIT: the stage would not fit)
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t adesc = umma_desc(a_addr + k * UK * 128, lbo, MN_SBO, MN_LAYOUT);
              const uint64_t adesc2 = umma_desc(a_addr + (128 / KC) * lbo + k * UK * 128, lbo, MN_SBO, MN_LAYOUT);
              const uint64_t bdesc = umma_desc(b_addr + k * UK * 128, lbo, MN_SBO, MN_LAYOUT);
              tc_mma<TF32>(dacc, adesc, bdesc, idesc, (uint32_t)((it | k) != 0));
              tc_mma<TF32>(dacc + (uint32_t)p.acc_cols, adesc2, bdesc, idesc, (uint32_t)((it | k) != 0));
            }
          } else {
            for (int k = 0; k < ksteps; ++k) {
              const uint32_t ao = A_MN ? (uint32_t)(k * UK * 128) : (uint32_t)(k * 32);
              const uint32_t bo = B_MN ? (uint32_t)(k * UK * 128) : (uint32_t)(k * 32);
              const uint64_t adesc = A_MN ? umma_desc(a_addr + ao, lbo, MN_SBO, MN_LAYOUT) : umma_desc_sw128(a_addr + ao, 16, 1024);
              const uint64_t bdesc = B_MN ? umma_desc(b_addr + bo, lbo, MN_SBO, MN_LAYOUT) : umma_desc_sw128(b_addr + bo, 16, 1024);
              // The tensor core's fp32 accumulate truncates: every MMA into an accumulator loses up to one ulp OF THE
              // ACCUMULATOR, towards zero, so a chain of n MMAs comes out short by ~n * 2^-24 (fc6, K = 25088: 2e-4).
              // With a second accumulator (nacc = 2) the split format keeps its two small correction terms apart from the
              // hi*hi chain (3x fewer truncations at full magnitude), the other formats alternate K steps (2x fewer).
              if (SPLIT) {  // hi*hi  +  (A_lo * B_hi + A_hi * B_lo)
                const uint64_t adesc_lo = A_MN ? umma_desc(a_addr + a_bytes + ao, lbo, MN_SBO, MN_LAYOUT)
                                               : umma_desc_sw128(a_addr + a_bytes + ao, 16, 1024);
                const uint64_t bdesc_lo = B_MN ? umma_desc(b_addr + b_bytes + bo, lbo, MN_SBO, MN_LAYOUT)
                                               : umma_desc_sw128(b_addr + b_bytes + bo, 16, 1024);
                const uint32_t dcorr = dacc + (uint32_t)((p.nacc - 1) * p.acc_cols);
                tc_mma<TF32>(dacc, adesc, bdesc, idesc, (uint32_t)((it | k) != 0));
                tc_mma<TF32>(dcorr, adesc_lo, bdesc, idesc, (uint32_t)(p.nacc == 1 || (it | k) != 0));
                tc_mma<TF32>(dcorr, adesc, bdesc_lo, idesc, 1u);
              } else {
                const int a = k & (p.nacc - 1);
                tc_mma<TF32>(dacc + (uint32_t)(a * p.acc_cols), adesc, bdesc, idesc, (uint32_t)(it != 0 || k >= p.nacc));
              }
            }
          }
          tc_commit(&empty[s]);  // frees the smem slot once these MMAs have read it
        }
        __syncwarp();
        if (++s == stages) s = 0, ph ^= 1u;
      }
      if (elect_one()) tc_commit(&accf[buf]);  // accumulator complete
      __syncwarp();
    }
  } else if (warp >= 2) {
    // =========================== epilogue ===========================
    constexpr int OUT_F32_ONLY = TF32 || MODE == 2;
    const bool f32_out = OUT_F32_ONLY || p.out_fp32;
    const int CW = f32_out ? 32 : 64;  // columns per 128-byte staging row
    constexpr int CWMAX = OUT_F32_ONLY ? 32 : 64;
    const int q4 = warp & 3;           // TMEM lane quarter this warp may read
    const int row = q4 * 32 + lane;
    const bool issuer = elect_one() && warp == 2;  // one fixed lane of warp 2 owns the bulk-store groups
    uint32_t local = 0, chunk_ctr = 0;

    // one 128 x 128-byte tile: registers -> swizzled staging rows -> TMA store / reduce-add.  Two staging buffers: the
    // store issued from a buffer two tiles ago must have read it before it is overwritten.
    // SWIZZLE_128B: 16-byte chunk j of row r lives at chunk (j ^ (r & 7)); conflict-free for a warp's 32 rows
    auto emit = [&](const uint4 (&q)[8], int nb, int m_row, const TileCoord& t, int plane) {
      uint8_t* sbuf = staging + (chunk_ctr & 1u) * STAGING_BYTES;
      ++chunk_ctr;
      if (issuer) bulk_wait_read<1>();
      named_bar_sync(1, 128);
      uint4* srow = reinterpret_cast<uint4*>(sbuf + row * 128);
#pragma unroll
      for (int j = 0; j < 8; ++j) srow[j ^ (row & 7)] = q[j];
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (issuer) {
        if (MODE == 2) tma_reduce_add_2d(&tmO, sbuf, nb, m_row);
        else if (p.im2col && SPLIT && !f32_out) tma_store_3d(&tmO, sbuf, nb, t.q_begin * 128, plane);  // {C, pixels, plane}
        else if (p.im2col) tma_store_2d(&tmO, sbuf, nb, t.q_begin * 128);                               // {C, pixels}
        else if (SPLIT && !f32_out) tma_store_5d(&tmO, sbuf, nb, t.x0, t.y0, t.b, plane);
        else tma_store_4d(&tmO, sbuf, nb, t.x0, t.y0, t.b);
        bulk_commit();
      }
    };

    for (uint32_t qi = 0;; ++qi) {
      const int slot = qi & (SQ - 1);
      mbar_wait(&sqf[slot], (qi / SQ) & 1u);
      const int tile = sq_tile[slot];
      __syncwarp();
      if (lane == 0) mbar_arrive(&sqe[slot]);
      if (tile < 0) break;
      const TileCoord t = decode_tile<MODE>(p, tile);
      if (t.n_iters == 0) continue;
      const uint32_t buf = p.nbuf == 2 ? (local & 1u) : 0u, aph = p.nbuf == 2 ? ((local >> 1) & 1u) : (local & 1u);
      ++local;
      bool ok = true;
      size_t orow = 0;  // output row index (pixel) for the mask / scale lookups
      int img = 0;
      if (MODE != 2 && p.im2col) {
        orow = (size_t)t.q_begin * 128 + row;  // the tile is a run of consecutive output pixels
        ok = (long long)orow < p.m_total;
        img = ok ? (int)(orow / (size_t)p.pix_per_image) : 0;
      } else if (MODE != 2) {
        const int ty = row / p.TW, tx_ = row - ty * p.TW;
        const int y = t.y0 + ty, x = t.x0 + tx_;
        ok = row < rows_a && y < p.H && x < p.W;
        orow = ((size_t)t.b * p.H + y) * p.W + x;
        img = ok ? (int)(orow / (size_t)p.pix_per_image) : 0;  // rows outside the image are clipped by the store
      }
      // dgrad: the ReLU-gate row of a chunk (128 bytes of the consumer activation) does not depend on the accumulator:
      // it is fetched one chunk ahead -- the first one before waiting for the MMAs -- so its global-memory latency hides
      // behind the TMEM load / staging / store of the previous chunk instead of stalling every chunk (dgrad ran 7-30 %
      // behind the forward pass of the same shape).  Not in the split format (254 registers already).
      constexpr bool PREF = MODE != 2 && !SPLIT;
      uint4 mrow[PREF ? 8 : 1];
      bool mrow_ok = false;  // mrow holds the whole row of the chunk about to be processed
      auto prefetch_mask = [&](int nb_) {
        mrow_ok = false;
        if (PREF && p.mask_ref && ok && nb_ + CW <= p.N) {
          const uint4* r4 = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.mask_ref) + orow * p.ld_mask + nb_);
#pragma unroll
          for (int j = 0; j < (PREF ? 8 : 1); ++j) mrow[j] = __ldg(r4 + j);
          mrow_ok = true;
        }
      };
      if (PREF) prefetch_mask(t.n0);
      mbar_wait(&accf[buf], aph);
      tc_fence_after();
      const uint32_t tbase = tmem + buf * (uint32_t)p.tmem_cols + ((uint32_t)(q4 * 32) << 16);

      const int n_chunks = (block_n + CW - 1) / CW;
      for (int hc = 0; hc < mp * n_chunks; ++hc) {
        const int h = hc / n_chunks, c = hc - h * n_chunks;  // h: 128-row sub-tile (MODE 2 with mpair = 2)
        const uint32_t tb = tbase + (uint32_t)(h * p.acc_cols);
        const int c0 = c * CW;
        const int nb = t.n0 + c0;
        const bool last = (h == mp - 1) && ((c == n_chunks - 1) || (nb + CW >= p.N));
        if (nb >= p.N) continue;  // uniform: the whole chunk lies in the column padding
        float f[CWMAX];
        {
          uint32_t v[32];
          const bool two = MODE != 2 && p.nacc == 2;  // second accumulator of the tile (see the MMA issuer)
          tmem_ld32(tb + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          if (two) {
            tmem_ld32(tb + (uint32_t)(p.acc_cols + c0), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
          }
          if (CWMAX == 64) {
            if (!f32_out) {
              tmem_ld32(tb + (uint32_t)(c0 + 32), v);
              tmem_ld_wait();
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) f[(32 + j) % CWMAX] = f32_out ? 0.f : __uint_as_float(v[j]);
            if (two && !f32_out) {
              tmem_ld32(tb + (uint32_t)(p.acc_cols + c0 + 32), v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) f[(32 + j) % CWMAX] += __uint_as_float(v[j]);
            }
          }
        }
        if (last) {  // every TMEM read of this tile is done: hand the accumulator buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acce[buf]);
        }
        const int nvalid = (p.N - nb) < CW ? (p.N - nb) : CW;
        if (MODE != 2) {
          // bias / Dropout2d scale rows: issue every load before the first use (32 dependent scalar loads per chunk
          // cost ~3.7k cycles in the first trace), 16-byte vectors when the row is whole and aligned
          const bool vec = (nvalid == CW) && p.vec_ok;
          if (p.bias) {
            float bv[CWMAX];
            load_row<CWMAX>(p.bias + nb, CW, nvalid, vec, bv);
#pragma unroll
            for (int j = 0; j < CWMAX; ++j) f[j] += bv[j];
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < CWMAX; ++j) f[j] = fmaxf(f[j], 0.f);
          }
          if (p.scale) {
            float sv[CWMAX];
            load_row<CWMAX>(p.scale + (size_t)img * p.scale_ld + nb, CW, nvalid, vec, sv);
#pragma unroll
            for (int j = 0; j < CWMAX; ++j) f[j] *= sv[j];
          }
          if (p.mask_ref && ok) {
            // ReLU gate: the sign of the stored activation (split: of its hi plane, which has the sign of the value)
            const T* ref = reinterpret_cast<const T*>(p.mask_ref) + orow * p.ld_mask + nb;
            if (nvalid == CW) {
              const uint4* r4 = reinterpret_cast<const uint4*>(ref);  // 128 contiguous bytes of this thread's row
              if (TF32) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const uint4 u = (PREF && mrow_ok) ? mrow[j % (PREF ? 8 : 1)] : __ldg(r4 + j);
                  if (!(__uint_as_float(u.x) > 0.f)) f[4 * j + 0] = 0.f;
                  if (!(__uint_as_float(u.y) > 0.f)) f[4 * j + 1] = 0.f;
                  if (!(__uint_as_float(u.z) > 0.f)) f[4 * j + 2] = 0.f;
                  if (!(__uint_as_float(u.w) > 0.f)) f[4 * j + 3] = 0.f;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const uint4 u = (PREF && mrow_ok) ? mrow[j % (PREF ? 8 : 1)] : __ldg(r4 + j);
                  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                    const uint32_t lo = w[e] & 0xFFFFu, hi = w[e] >> 16;
                    if ((lo & 0x8000u) || (lo & 0x7FFFu) == 0) f[(8 * j + 2 * e) % CWMAX] = 0.f;
                    if ((hi & 0x8000u) || (hi & 0x7FFFu) == 0) f[(8 * j + 2 * e + 1) % CWMAX] = 0.f;
                  }
                }
              }
            } else {
              // (compile-time indices only: a runtime index would push f[] into local memory)
#pragma unroll
              for (int j = 0; j < CWMAX; ++j)
                if (j < nvalid && !(load_as_float<T>(ref + j) > 0.f)) f[j] = 0.f;
            }
          }
          if (PREF && p.mask_ref && hc + 1 < mp * n_chunks) prefetch_mask(nb + CW);  // next chunk's gate row
        }
        // ---- registers -> staging -> TMA store ----
        uint4 q[8];
        if (f32_out) {
          if (TF32 && MODE != 2 && !p.out_fp32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = to_tf32(f[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            q[j] = make_uint4(__float_as_uint(f[4 * j]), __float_as_uint(f[4 * j + 1]), __float_as_uint(f[4 * j + 2]),
                              __float_as_uint(f[4 * j + 3]));
          emit(q, nb, t.m0 + h * 128, t, 0);
        } else {
          // bf16 output; SPLIT: the hi plane now, the lo plane (the rounding residuals) right after it
          float res[CWMAX];  // SPLIT only: what the hi plane's rounding left over
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h2[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i0 = (8 * j + 2 * e) % CWMAX, i1 = (8 * j + 2 * e + 1) % CWMAX;
              h2[e] = __floats2bfloat162_rn(f[i0], f[i1]);
              if (SPLIT) {
                res[i0] = f[i0] - __low2float(h2[e]);
                res[i1] = f[i1] - __high2float(h2[e]);
              } else if (MODE != 2 && p.col_sum) {  // the column sums must see exactly the values that are stored
                f[i0] = __low2float(h2[e]), f[i1] = __high2float(h2[e]);
              }
            }
            q[j].x = *reinterpret_cast<uint32_t*>(&h2[0]);
            q[j].y = *reinterpret_cast<uint32_t*>(&h2[1]);
            q[j].z = *reinterpret_cast<uint32_t*>(&h2[2]);
            q[j].w = *reinterpret_cast<uint32_t*>(&h2[3]);
          }
          emit(q, nb, 0, t, 0);
          if (SPLIT) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 l2[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                l2[e] = __floats2bfloat162_rn(res[(8 * j + 2 * e) % CWMAX], res[(8 * j + 2 * e + 1) % CWMAX]);
              q[j].x = *reinterpret_cast<uint32_t*>(&l2[0]);
              q[j].y = *reinterpret_cast<uint32_t*>(&l2[1]);
              q[j].z = *reinterpret_cast<uint32_t*>(&l2[2]);
              q[j].w = *reinterpret_cast<uint32_t*>(&l2[3]);
            }
            emit(q, nb, 0, t, 1);
          }
        }
        if (MODE != 2 && p.col_sum) {
          // bias gradient of the layer that produced this dY: column sums of the stored tile (split: of hi + lo up to the
          // 2^-17 of the second rounding).  Butterfly transpose-reduce: after 31 shuffles lane l holds the sum over the
          // warp's 32 rows of column l; one RED per (warp, column).
          const bool row_live = ok;
#pragma unroll
          for (int half = 0; half < CWMAX / 32; ++half) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = row_live ? f[half * 32 + j] : 0.f;
#pragma unroll
            for (int sft = 16; sft >= 1; sft >>= 1) {
              const bool up = (lane & sft) != 0;
#pragma unroll
              for (int i = 0; i < sft; ++i) {
                const float send = up ? v[i] : v[i + sft];
                const float keep = up ? v[i + sft] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
              }
            }
            const int col = nb + half * 32 + lane;
            if (half * 32 < CW && col < p.N) {
              if (p.cs_local) atomicAdd(cs_smem + col, v[0]);  // one N tile: col < block_n <= 256
              else atomicAdd(p.col_sum + col, v[0]);
            }
          }
        }
      }
    }
    if (MODE != 2 && p.cs_local) {
      named_bar_sync(2, 128);  // every epilogue warp has added its last tile (barrier 1 belongs to emit())
      for (int j = (int)threadIdx.x - 64; j < p.N; j += 128) {
        const float v = cs_smem[j];
        if (v != 0.f) atomicAdd(p.col_sum + j, v);
      }
    }
    if (issuer) bulk_wait<0>();  // all stores have landed before the CTA (and its shared memory) goes away
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, (uint32_t)(p.nbuf * p.tmem_cols));
  if (threadIdx.x == 0) {
    // the last CTA to leave re-arms the work counter for the next launch that uses this slot
    __threadfence();
    if (atomicAdd(p.sched + 1, 1u) == gridDim.x - 1) {
      p.sched[0] = 0u;
      p.sched[1] = 0u;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// dims/box innermost first; strides in elements for dims 1..rank-1
// dtype here is the ELEMENT type of the mapped tensor (SZN_F32 / SZN_BF16)
// Encoded maps are cached, keyed by (base pointer, element type, rank, dims, strides, box, swizzle): a training loop
// re-uses the same buffers step after step (the caching allocator hands the same blocks back, weights never move), so
// after the first step every launch finds its three maps ready instead of making three driver calls.
struct TmapKey {
  const void* base;
  int dtype, rank, mn;
  long long dims[5], strides[5];
  int box[5];
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapSlot {
  TmapKey key;
  CUtensorMap map;
  bool used;
};
constexpr int TMAP_SLOTS = 1024;  // direct-mapped

static int make_tmap(CUtensorMap* m, int dtype, const void* base, int rank, const long long* dims,
                     const long long* strides_elems, const int* box, bool mn_major = false) {
  static TmapSlot* cache = nullptr;  // per process; the map embeds the pointer, which is device specific by construction
  if (!cache) cache = static_cast<TmapSlot*>(calloc(TMAP_SLOTS, sizeof(TmapSlot)));
  TmapKey key;
  memset(&key, 0, sizeof key);
  key.base = base, key.dtype = dtype, key.rank = rank, key.mn = mn_major ? 1 : 0;
  for (int i = 0; i < rank; ++i) key.dims[i] = dims[i], key.strides[i] = i ? strides_elems[i] : 1, key.box[i] = box[i];
  size_t h = 1469598103934665603ull;
  for (size_t i = 0; i < sizeof key; ++i) h = (h ^ reinterpret_cast<const unsigned char*>(&key)[i]) * 1099511628211ull;
  TmapSlot* slot = cache ? &cache[h % TMAP_SLOTS] : nullptr;
  if (slot && slot->used && slot->key == key) {
    *m = slot->map;
    return 0;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(SZN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const int es = dtype == SZN_BF16 ? 2 : 4;
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], el[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
    el[i] = 1;
    if (i > 0) gs[i - 1] = (cuuint64_t)strides_elems[i] * es;
  }
  for (int i = 0; i + 1 < rank; ++i)
    if (gs[i] % 16) return set_error(SZN_ERR_ARG, "TMA global stride not a multiple of 16 bytes");
  if (reinterpret_cast<uintptr_t>(base) % 16) return set_error(SZN_ERR_ARG, "TMA base not 16-byte aligned");
  CUresult r = enc(m, dtype == SZN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                   (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   // operands consumed MN-major in tf32 need 32-byte swizzle atoms (see umma_desc)
                   (mn_major && dtype == SZN_F32) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[160];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeTiled failed (%d) rank %d dims %lld %lld box %d %d", (int)r, rank,
             dims[0], dims[1], box[0], box[1]);
    return set_error(SZN_ERR_CUDA, msg);
  }
  if (slot) {
    slot->key = key;
    slot->map = *m;
    slot->used = true;
  }
  return 0;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// im2col-mode map of an NHWC tensor [B][H][W][pitch] (pitch >= C elements per pixel) for a conv with symmetric padding `pad`
// and an R x S filter: a load lands 128 consecutive output pixels x `kc` channels at a tap offset (tma_load_im2col_4d).
// Corner convention (CUDA driver API / CUTLASS copy_traits_sm90_im2col): lower = -pad, upper = pad - (filter - 1).
static int make_tmap_im2col(CUtensorMap* m, int dtype, const void* base, long long C, long long pitch, int W, int H, int B,
                            int R, int S, int pad, int kc) {
  static TmapSlot* cache = nullptr;
  if (!cache) cache = static_cast<TmapSlot*>(calloc(TMAP_SLOTS, sizeof(TmapSlot)));
  TmapKey key;
  memset(&key, 0, sizeof key);
  key.base = base, key.dtype = dtype, key.rank = 4, key.mn = 2;  // mn = 2 marks the im2col flavour
  key.dims[0] = C, key.dims[1] = W, key.dims[2] = H, key.dims[3] = B, key.dims[4] = pitch;
  key.box[0] = kc, key.box[1] = R, key.box[2] = S, key.box[3] = pad;
  size_t h = 1469598103934665603ull;
  for (size_t i = 0; i < sizeof key; ++i) h = (h ^ reinterpret_cast<const unsigned char*>(&key)[i]) * 1099511628211ull;
  TmapSlot* slot = cache ? &cache[h % TMAP_SLOTS] : nullptr;
  if (slot && slot->used && slot->key == key) {
    *m = slot->map;
    return 0;
  }
  static EncodeIm2colFn enc = nullptr;
  if (!enc) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return set_error(SZN_ERR_CUDA, "cuTensorMapEncodeIm2col entry point not available");
    enc = reinterpret_cast<EncodeIm2colFn>(ptr);
  }
  const int es = dtype == SZN_BF16 ? 2 : 4;
  cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gs[3] = {(cuuint64_t)pitch * es, (cuuint64_t)W * pitch * es, (cuuint64_t)H * W * pitch * es};
  int lower[2] = {-pad, -pad}, upper[2] = {pad - (S - 1), pad - (R - 1)};  // (W, H)
  cuuint32_t el[4] = {1, 1, 1, 1};
  if (gs[0] % 16 || reinterpret_cast<uintptr_t>(base) % 16) return set_error(SZN_ERR_ARG, "im2col TMA: base / pixel pitch not 16-byte aligned");
  CUresult r = enc(m, dtype == SZN_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                   const_cast<void*>(base), gd, gs, lower, upper, (cuuint32_t)kc, 128, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[200];
    snprintf(msg, sizeof msg, "cuTensorMapEncodeIm2col failed (%d): C %lld pitch %lld W %d H %d B %d R %d S %d pad %d", (int)r, C,
             pitch, W, H, B, R, S, pad);
    return set_error(SZN_ERR_CUDA, msg);
  }
  {
    // drivers up to CUDA 13.1 mis-encode im2col maps of tensors smaller than 128 KB (the workaround CUTLASS ships in
    // copy_traits_sm90_im2col.hpp: clear bit 21 of the second descriptor word)
    static int drv = -1;
    if (drv < 0 && cudaDriverGetVersion(&drv) != cudaSuccess) drv = 0;
    if (drv <= 13010 && (unsigned long long)B * H * W * pitch * es < 131072ull) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  }
  if (slot) {
    slot->key = key;
    slot->map = *m;
    slot->used = true;
  }
  return 0;
}

static int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// pick TW x TH (TW*TH <= max_rows) minimising the number of tiles over a W x H image
static void pick_tile(int W, int H, int max_rows, int* TW, int* TH) {
  long long best = -1;
  int bw = max_rows, bh = 1;
  for (int tw = 1; tw <= max_rows && tw <= 256; ++tw) {
    int th = max_rows / tw;
    if (th < 1) break;
    if (th > 256) th = 256;
    // only consider the widest useful tiles: powers of two and exact widths
    const bool pow2 = (tw & (tw - 1)) == 0;
    if (!pow2 && tw != W && tw != (W + 1) / 2) continue;
    const long long cnt = (long long)ceil_div(W, tw) * ceil_div(H, th);
    // prefer fewer tiles, then wider tiles (longer contiguous runs)
    if (best < 0 || cnt < best || (cnt == best && tw > bw)) {
      best = cnt;
      bw = tw;
      bh = th;
    }
  }
  if (bh > H) bh = H;  // never ask for more rows than exist (keeps the box tight)
  *TW = bw;
  *TH = bh;
}

static int tmem_cols_for(int n) { return n <= 32 ? 32 : n <= 64 ? 64 : n <= 128 ? 128 : 256; }

static int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// work counters of the dynamic tile scheduler: {next, done} pairs, zero at module load and re-armed by the kernel itself.
// Launches take the slots round-robin, so kernels that overlap on different streams do not share a counter.
constexpr int SCHED_SLOTS = 256;
__device__ unsigned int g_sched[2 * SCHED_SLOTS];

static unsigned int* sched_slot() {
  static unsigned int* base[64] = {nullptr};
  static unsigned int seq = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (!base[dev] && cudaGetSymbolAddress(reinterpret_cast<void**>(&base[dev]), g_sched) != cudaSuccess) return nullptr;
  return base[dev] + 2 * (seq++ % SCHED_SLOTS);
}

template <typename T, int MODE, bool SPLIT>
static int launch(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, UmmaParams& p, long long tiles,
                  cudaStream_t st) {
  if (MODE != 2 || p.mpair < 1 || SPLIT) p.mpair = 1;
  const int stage_bytes = (SPLIT ? 2 : 1) * (128 * 128 * p.mpair + p.block_n * 128);
  const int fixed = 2 * 128 * 128 /* epilogue staging */ + 1024 /* alignment */ + 384 /* barriers, tile queue */ +
                    1024 /* per-CTA column sums */;
  int stages = (227 * 1024 - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) return set_error(SZN_ERR_UNSUPPORTED, "conv: tile does not leave two pipeline stages");
  p.stages = stages;
  p.acc_cols = tmem_cols_for(p.block_n);
  // second accumulator (shorter truncating fp32 chains, see the MMA issuer): whenever TMEM still holds two buffers of
  // two; for 256-column tiles only when the main loop is so long (fc6: 392 stages) that giving up the overlap of a
  // tile's epilogue with the next main loop costs nothing measurable
  p.nacc = 1;
  if (MODE == 0) {
    const int n_iters = p.R * p.S * p.kchunks;
    if (4 * p.acc_cols <= 512 && (SPLIT || n_iters >= 64)) p.nacc = 2;
    else if (2 * p.acc_cols <= 512 && n_iters >= 256) p.nacc = 2;
  }
  p.tmem_cols = p.acc_cols * p.mpair * p.nacc;  // per accumulator buffer
  p.nbuf = 2 * p.tmem_cols <= 512 ? 2 : 1;
  p.total_tiles = (int)tiles;
  {
    static int st = -1;
    if (st < 0) st = getenv("SZN_STATIC_TILES") ? 1 : 0;
    p.static_tiles = st;
  }
  p.sched = sched_slot();
  if (!p.sched) return set_error(SZN_ERR_CUDA, "conv: no work counter (cudaGetSymbolAddress failed)");
  const size_t smem = (size_t)stages * stage_bytes + fixed;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(umma_conv_kernel<T, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return set_error(SZN_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  // persistent: one CTA per SM walks the tile list
  const unsigned grid = (unsigned)(tiles < num_sms() ? tiles : num_sms());
  umma_conv_kernel<T, MODE, SPLIT><<<grid, 192, smem, st>>>(a, b, o, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SZN_ERR_CUDA, cudaGetErrorString(e));
  count_launch();
  return 0;
}

template <int MODE>
static int launch_dtype(int dtype, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, UmmaParams& p,
                        long long tiles, cudaStream_t st) {
  if (dtype == SZN_BF16) return launch<__nv_bfloat16, MODE, false>(a, b, o, p, tiles, st);
  if (dtype == SZN_F32X3) return launch<__nv_bfloat16, MODE, true>(a, b, o, p, tiles, st);
  return launch<float, MODE, false>(a, b, o, p, tiles, st);
}

// shrink the N tile while the launch would leave SMs idle (fewer CTAs than SMs), keeping N % block_n == 0
static int fill_sms(int block_n, int N, int gran, long long m_tiles) {
  while (block_n > 64 && m_tiles * ((N + block_n - 1) / block_n) < 148) {
    int nb = block_n / 2;
    if (nb % gran || N % nb) break;
    block_n = nb;
  }
  return block_n;
}

static int pick_block_n(int N, int gran, int maxn) {
  // largest tile <= maxn (multiple of gran) that minimises padded columns
  int best = gran, best_waste = 1 << 30;
  for (int bn = gran; bn <= maxn; bn += gran) {
    const int tiles = ceil_div(N, bn);
    const int waste = tiles * bn - N;
    if (waste < best_waste || (waste == best_waste && bn > best)) {
      best = bn;
      best_waste = waste;
    }
  }
  return best;
}

}  // namespace szn

using namespace szn;

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
// shared by the forward conv and the data gradient (which is a forward conv of dY with transposed, flipped weights)
static int conv_gemm(int dtype, const void* x, long long ldx, const void* wt, const float* bias, void* y, int B, int H,
                     int W, int Cin, int Cout, int R, int S, int pad, int relu, const float* scale, int scale_ld,
                     int out_fp32, long long ldo, const void* mask_ref, float* col_sum, void* stream) {
  if (dtype != SZN_F32 && dtype != SZN_BF16 && dtype != SZN_F32X3) return set_error(SZN_ERR_ARG, "conv: bad dtype");
  const bool split = dtype == SZN_F32X3;
  const int edt = dtype == SZN_F32 ? SZN_F32 : SZN_BF16;  // element type of one operand plane
  const int KC = edt == SZN_BF16 ? 64 : 32;
  if (Cin % KC && !(R == 1 && S == 1)) return set_error(SZN_ERR_ARG, "conv: channels per tap must be a multiple of 128 bytes");
  if (split && (Cin % 8 || ldx % 8 || ldo % 8)) return set_error(SZN_ERR_ARG, "conv (split): planes must be 16-byte aligned");
  int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  if (Ho <= 0 || Wo <= 0) return set_error(SZN_ERR_ARG, "conv: empty output");
  UmmaParams p{};
  p.pix_per_image = Ho * Wo;
  int Bq = B, Hq = H, Wq = W;
  if (R == 1 && S == 1 && pad == 0) {  // 1x1: flatten every pixel of the batch into one row of pixels
    Wq = B * H * W, Hq = 1, Bq = 1, Ho = 1, Wo = Wq;
  }
  // k > 1: M tiles are runs of 128 consecutive output pixels fetched by im2col-mode TMA (no rectangular tiles, so no
  // padded tile rows: rectangles over-covered the 710 / 355 / 178 / 89 / 45-pixel maps by 8-14 %).  1x1: the flattened
  // pixel row below is already that.  SZN_NO_IM2COL=1 keeps the rectangular tiles (A/B switch).
  static int no_im2col = -1;
  if (no_im2col < 0) no_im2col = getenv("SZN_NO_IM2COL") ? 1 : 0;
  const bool im2col = !(R == 1 && S == 1 && pad == 0) && !no_im2col;
  p.im2col = im2col ? 1 : 0;
  p.m_total = (long long)Bq * Ho * Wo;
  if (im2col) {
    p.TW = 128, p.TH = 1;
    p.tiles_x = ceil_div(p.m_total, 128), p.tiles_y = 1, p.B = 1;
  } else {
    pick_tile(Wo, Ho, 128, &p.TW, &p.TH);
    p.tiles_x = ceil_div(Wo, p.TW), p.tiles_y = ceil_div(Ho, p.TH), p.B = Bq;
  }
  const long long m_tiles = (long long)p.tiles_x * p.tiles_y * p.B;
  p.H = Ho, p.W = Wo, p.R = R, p.S = S, p.pad = pad, p.Ck = Cin, p.kchunks = ceil_div(Cin, KC);
  p.N = Cout;
  const int out_f32 = (dtype == SZN_F32 || out_fp32) ? 1 : 0;
  const int ngran = out_f32 ? 32 : 64;  // the epilogue moves 128-byte output rows
  p.block_n = pick_block_n(Cout, ngran, Cout >= 256 ? 256 : 128);
  p.block_n = fill_sms(p.block_n, Cout, ngran, m_tiles);
  p.n_tiles = ceil_div(Cout, p.block_n);
  p.ldo = ldo, p.bias = bias, p.scale = scale, p.scale_ld = scale_ld, p.relu = relu, p.out_fp32 = out_fp32;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(scale)) & 15) == 0 && scale_ld % 4 == 0;
  p.mask_ref = mask_ref;
  p.ld_mask = split ? 2 * ldo : ldo;
  p.col_sum = col_sum;
  p.cs_local = (col_sum && p.n_tiles == 1 && p.block_n <= 256 && !env_flag("SZN_COLSUM_GLOBAL", 0)) ? 1 : 0;
  CUtensorMap ta, tb, to;
  {
    // split: a pixel row is [hi | lo], i.e. twice the row pitch with the plane as one more (outermost) dimension
    const long long opitch = (split && !out_f32) ? 2 * ldo : ldo, xpitch = split ? 2 * ldx : ldx;
    if (im2col) {
      // output: the flat [pixels][channels] matrix (+ the plane as a third dimension in the split format)
      long long od[3] = {Cout, p.m_total, 2}, os[3] = {1, opitch, ldo};
      int obx[3] = {out_f32 ? 32 : 64, 128, 1};
      if (int e = make_tmap(&to, out_f32 ? SZN_F32 : SZN_BF16, y, (split && !out_f32) ? 3 : 2, od, os, obx)) return e;
      // input: NHWC, a pixel row holds Cin channels (split: [hi | lo] = 2 Cin bf16)
      p.a_lo_ch = (int)ldx;
      if (int e = make_tmap_im2col(&ta, edt, x, split ? ldx + Cin : Cin, xpitch, Wq, Hq, Bq, R, S, pad, KC)) return e;
    } else {
      long long od[5] = {Cout, Wo, Ho, Bq, 2}, os[5] = {1, opitch, (long long)Wo * opitch, (long long)Ho * Wo * opitch, ldo};
      int obx[5] = {out_f32 ? 32 : 64, p.TW, p.TH, 1, 1};
      if (int e = make_tmap(&to, out_f32 ? SZN_F32 : SZN_BF16, y, (split && !out_f32) ? 5 : 4, od, os, obx)) return e;
      long long d[5] = {Cin, Wq, Hq, Bq, 2}, s[5] = {1, xpitch, (long long)Wq * xpitch, (long long)Hq * Wq * xpitch, ldx};
      int bx[5] = {KC, p.TW, p.TH, 1, 1};
      if (int e = make_tmap(&ta, edt, x, split ? 5 : 4, d, s, bx)) return e;
    }
    long long K = (long long)R * S * Cin;
    long long d2[3] = {K, Cout, 2}, s2[3] = {1, K, (long long)Cout * K};  // split: hi plane, then lo plane
    int bx2[3] = {KC, p.block_n, 1};
    if (int e = make_tmap(&tb, edt, wt, split ? 3 : 2, d2, s2, bx2)) return e;
  }
  const long long tiles = m_tiles * p.n_tiles;
  {
    // Which operand should concurrently running CTAs share through L2?  By default the pixel tile (N tiles fastest).
    // When the weights are far larger than L2 while the activations fit (fc6's data gradient: 411 MB of weights, 38 MB
    // of dY), N-fastest makes every M tile re-stream all weights from HBM; M-fastest reads each weight slice once.
    const double es = dtype == SZN_BF16 ? 2.0 : 4.0;
    const double w_bytes = (double)R * S * Cin * Cout * es, a_bytes = (double)B * H * W * Cin * es;
    p.m_fast = (w_bytes > 96e6 && a_bytes < 64e6 && p.n_tiles > 1) ? 1 : 0;
  }
  return launch_dtype<0>(dtype, ta, tb, to, p, tiles, (cudaStream_t)stream);
}

extern "C" int szn_conv_fwd(int dtype, const void* x, const void* wt, const float* bias, void* y, int B, int H, int W,
                            int Cin, int Cout, int R, int S, int pad, int relu, const float* scale, int scale_ld,
                            int out_fp32, long long ldo, void* stream) {
  return conv_gemm(dtype, x, Cin, wt, bias, y, B, H, W, Cin, Cout, R, S, pad, relu, scale, scale_ld, out_fp32, ldo,
                   nullptr, nullptr, stream);
}

// dx[B,H,W,Cin] (the conv input's gradient) from dy[B,Ho,Wo,Cout]:
//   dx[p, ci] = sum_{r,s,co} dy[p + pad - (r,s), co] * w[co, ci, r, s]
//             = sum_{r',s',co} dy[p + (r',s') - (R-1-pad), co] * wt_d[ci][r',s'][co],   wt_d[ci][r'][s'][co] = w[co][ci][R-1-r'][S-1-s']
// i.e. a forward convolution of dy with padding R-1-pad and the transposed, flipped weights that szn_pack_weight_dgrad
// writes, so it runs on the K-major kernel path with one weight box per stage.  Epilogue: optional ReLU gate by
// `relu_ref` (same shape as dx) and per-(image, channel) multiplier `scale`.
extern "C" int szn_conv_dgrad(int dtype, const void* dy, const void* wt_dgrad, void* dx, int B, int H, int W, int Cin,
                              int Cout, int R, int S, int pad, const void* relu_ref, const float* scale, int scale_ld,
                              long long ld_dy, float* dx_col_sum, void* stream) {
  const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  return conv_gemm(dtype, dy, ld_dy, wt_dgrad, nullptr, dx, B, Ho, Wo, Cout, Cin, R, S, R - 1 - pad, 0, scale, scale_ld,
                   0, Cin, relu_ref, dx_col_sum, stream);
}

// dw[Cout][R*S*Cin] (fp32, ACCUMULATED into: the caller zeroes it) from x[B,H,W,Cin] and dy[B,Ho,Wo,Cout]
// K-chunk shape of the wgrad: TW x TH pixels with TW*TH <= max_rows and a multiple of the UMMA K step, minimising the
// padded pixel count over a W x H image (ties: wider)
static void pick_tile_k(int W, int H, int max_rows, int step, int* TW, int* TH) {
  long long best = -1;
  int bw = step, bh = 1;
  for (int tw = 1; tw <= max_rows; ++tw) {
    for (int th = max_rows / tw; th >= 1; --th) {
      if ((tw * th) % step) continue;
      const long long cost = (long long)ceil_div(W, tw) * ceil_div(H, th) * tw * th;
      if (best < 0 || cost < best || (cost == best && tw > bw)) best = cost, bw = tw, bh = th;
      break;  // smaller th for the same tw only helps through `cost`, examined via other tw
    }
  }
  *TW = bw;
  *TH = bh;
}

// split-K work items per CTA of the weight gradient.  1 (default) is fastest on an otherwise idle GPU (measured 5.30 / 5.38 /
// 5.46 / 5.54 ms per step for 1 / 2 / 3 / 4: every item ends with a reduce-add pass over its output tile); data-parallel
// runs use 3, so that a CTA whose SM is busy with NCCL's kernels costs a third of a tile list, not a whole one.
static int g_wgrad_waves = 0;
static int wgrad_waves() {
  if (!g_wgrad_waves) {
    const char* e = getenv("SZN_WGRAD_WAVES");
    g_wgrad_waves = e ? atoi(e) : 1;
    if (g_wgrad_waves < 1) g_wgrad_waves = 1;
  }
  return g_wgrad_waves;
}
extern "C" int szn_set_wgrad_waves(int waves) {
  if (waves < 1 || waves > 16) return set_error(SZN_ERR_ARG, "szn_set_wgrad_waves: 1..16");
  g_wgrad_waves = waves;
  return 0;
}

extern "C" int szn_conv_wgrad(int dtype, const void* x, const void* dy, float* dw, int B, int H, int W, int Cin,
                              int Cout, int R, int S, int pad, long long ld_dy, void* stream) {
  if (dtype != SZN_F32 && dtype != SZN_BF16 && dtype != SZN_F32X3) return set_error(SZN_ERR_ARG, "szn_conv_wgrad: bad dtype");
  const bool split = dtype == SZN_F32X3;
  const int edt = dtype == SZN_F32 ? SZN_F32 : SZN_BF16;
  const int KC = edt == SZN_BF16 ? 64 : 32, UK = KC / 4;
  if (Cin % KC || Cout % KC || ld_dy % KC)
    return set_error(SZN_ERR_ARG, "szn_conv_wgrad: Cin, Cout and ld_dy must be multiples of 128 bytes");
  int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  int Bq = B, Hq = H, Wq = W;
  if (R == 1 && S == 1 && pad == 0) Wq = B * H * W, Hq = 1, Bq = 1, Ho = 1, Wo = Wq;
  UmmaParams p{};
  pick_tile_k(Wo, Ho, KC, UK, &p.TW, &p.TH);
  p.ksteps = p.TW * p.TH / UK;
  p.tiles_x = ceil_div(Wo, p.TW), p.tiles_y = ceil_div(Ho, p.TH), p.B = Bq;
  p.R = R, p.S = S, p.pad = pad;
  p.M = Cout, p.Cin = Cin;
  p.N = R * S * Cin;
  // N tile: a divisor of Cin (one tap, one box of block_n/KC channel groups) or a multiple of it (m whole taps, m boxes)
  const int taps = R * S;
  if (Cin >= 256) {
    p.block_n = 256;
    while (Cin % p.block_n) p.block_n -= KC;
    p.gpt = p.block_n / KC, p.b_boxes = 1;
  } else {
    int best_m = 1;
    double best_cost = 1e30;
    for (int m = 1; m * Cin <= 256 && m <= 4; ++m) {
      const double cost = (double)ceil_div(taps, m) * m * (m * Cin < 256 ? 1.15 : 1.0);  // narrow tiles are smem-bound
      if (cost <= best_cost) best_cost = cost, best_m = m;
    }
    p.block_n = best_m * Cin, p.gpt = Cin / KC, p.b_boxes = best_m;
  }
  p.n_tiles = ceil_div(p.N, p.block_n);
  // two 128-row output sub-tiles per work item when Cout allows: they share every B (x) stage, so the bytes fetched per
  // MMA cycle drop from 96 to 64 -- the wgrad streams both operands once and is bound by load latency x ring depth
  {
    static int no_pair = -1;
    if (no_pair < 0) no_pair = getenv("SZN_NO_MPAIR") ? 1 : 0;
    p.mpair = (!no_pair && !split && Cout >= 256 && p.block_n == 256) ? 2 : 1;
  }
  p.m_tiles = ceil_div(Cout, 128 * p.mpair);
  const long long total_q = (long long)p.tiles_x * p.tiles_y * Bq;
  const long long tiles = (long long)p.n_tiles * p.m_tiles;
  const int waves = wgrad_waves();
  // `waves` work items per CTA, rounded DOWN so that no CTA gets one item more than the others (rounding up gave
  // e.g. 297 items for 148 CTAs: one CTA with 3 items set the kernel's duration)
  long long splits = ((long long)waves * num_sms()) / tiles;
  if (splits > total_q / 4) splits = total_q / 4;    // at least 4 chunks per split
  if (splits < 1) splits = 1;
  p.splits = (int)splits;
  p.ldo = p.N;
  CUtensorMap ta, tb, to;
  {
    long long od[2] = {p.N, Cout}, os[2] = {1, p.N};
    int obx[2] = {32, 128};
    if (int e = make_tmap(&to, SZN_F32, dw, 2, od, os, obx)) return e;
    // 5-D views {KC channels, W, H, B, channel group}: one box fetches several 128-byte channel groups of a pixel patch.
    // split: a pixel row is [hi | lo]; the lo plane's groups simply follow the hi plane's at a_lo_goff / b_lo_goff.
    const long long ypitch = split ? 2 * ld_dy : ld_dy, xpitch = split ? 2 * (long long)Cin : Cin;
    p.a_lo_goff = (int)(ld_dy / KC), p.b_lo_goff = Cin / KC;
    long long d[5] = {KC, Wo, Ho, Bq, split ? p.a_lo_goff + Cout / KC : Cout / KC};
    long long s[5] = {1, ypitch, (long long)Wo * ypitch, (long long)Ho * Wo * ypitch, KC};
    int bx[5] = {KC, p.TW, p.TH, 1, (128 / KC) * p.mpair};
    if (int e = make_tmap(&ta, edt, dy, 5, d, s, bx, true)) return e;
    long long d2[5] = {KC, Wq, Hq, Bq, (split ? 2 : 1) * Cin / KC};
    long long s2[5] = {1, xpitch, (long long)Wq * xpitch, (long long)Hq * Wq * xpitch, KC};
    int bx2[5] = {KC, p.TW, p.TH, 1, p.gpt};
    if (int e = make_tmap(&tb, edt, x, 5, d2, s2, bx2, true)) return e;
  }
  const long long grid = tiles * p.splits;
  return launch_dtype<2>(dtype, ta, tb, to, p, grid, (cudaStream_t)stream);
}
