// infer_lbl (utils.py:159-185) on tcgen05 tensor cores: labels[p] = argmax_c <s_p, e_c> / |e_c|  (|e_c| == 0 -> 1;
// the per-pixel norm |s_p| is a positive factor common to every class and cannot change the arg-max).
//
// The contraction [B*H*W, D] x [D, C] runs as a TF32 GEMM with error compensation ("3xTF32"): every fp32 operand is
// split into hi = its upper 19 bits and lo = the exact remainder, and D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo, which
// leaves a relative error of ~2^-22 -- the same order as the fp32 summation-order noise of any other implementation, so
// labels can differ from the fp32 oracle only where its own top-2 margin is below that noise.
//
//   A = score, NCHW fp32: pixels are contiguous per channel, i.e. the GEMM's M-major ("MN-major") layout.  One 4-D TMA
//       box {32 px, 32 channels, 1 image, 4 pixel groups} lands a 128-pixel x 32-channel tile (SWIZZLE_128B_BASE32B).
//       Four "split" warps then rewrite it in place as hi and write lo to a second buffer.
//   B = class table [C][D] (K-major), pre-split into hi / lo, zero-padded to [Cpad][Dpad], by szn_embed_argmax.
//   D = three accumulators in TMEM (hh, lh, hl: independent MMA chains; for 128 < C <= 256 two: hh and hl + lh),
//       summed by the epilogue, which scales by 1/|e_c|, takes the first maximum over c < C and writes the int64 label.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2-5 operand split, 6-9 epilogue.  Persistent over 128-pixel tiles.
#include "szn_internal.h"
#include "szn_ptx.cuh"

namespace szn {

struct ArgmaxParams {
  int D, C, Cpad, B;
  long long hw;
  int tiles_per_img, total_tiles, kchunks, stages, acc_cols, nbuf, nacc;
  const float* inv_en;  // [Cpad] 1 / |e_c| (1 for zero rows)
  long long* labels;
};

__global__ void table_split_kernel(const float* __restrict__ table, int C, int D, int Cpad, int Dpad,
                                   float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ inv_en) {
  const int c = blockIdx.x;
  __shared__ float red[128];
  float ss = 0.f;
  for (int d = threadIdx.x; d < Dpad; d += blockDim.x) {
    const float v = (c < C && d < D) ? table[(long long)c * D + d] : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    hi[(long long)c * Dpad + d] = h;
    lo[(long long)c * Dpad + d] = v - h;
    ss = fmaf(v, v, ss);
  }
  red[threadIdx.x] = ss;
  __syncthreads();
  for (int o = 64; o; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float n = sqrtf(red[0]);
    inv_en[c] = n == 0.f ? 1.f : 1.f / n;
  }
}

__global__ void __launch_bounds__(320, 1)
embed_argmax_tc_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmHi,
                       const __grid_constant__ CUtensorMap tmLo, const ArgmaxParams p) {
  constexpr int A_BYTES = 128 * 128;  // 4 pixel groups x 32 channels x 128 B
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int b_bytes = p.Cpad * 128;
  const int stage_bytes = 2 * A_BYTES + 2 * b_bytes;  // A_hi | A_lo | B_hi | B_lo
  const int stages = p.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + stages * stage_bytes);
  uint64_t* split = full + 8;   // A_hi / A_lo of the stage are ready for the MMAs
  uint64_t* empty = split + 8;
  uint64_t* accf = empty + 8;
  uint64_t* acce = accf + 2;
  uint32_t* tptr = reinterpret_cast<uint32_t*>(acce + 2);
  const int warp = warp_idx(), lane = threadIdx.x & 31;  // provably warp-uniform (see elect_one)

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmHi);
    tma_prefetch_desc(&tmLo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&split[i], 128);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&accf[i], 1);
      mbar_init(&acce[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  const uint32_t buf_cols = (uint32_t)(p.nacc * p.acc_cols);  // hh | hl | lh, or (Cpad = 256) hh | hl + lh

  if (warp == 0) {
    // ---------------- TMA producer (whole warp loops, one elected lane issues) ----------------
    int s = 0;
    uint32_t ph = 0;
    const uint32_t tx = (uint32_t)(A_BYTES + 2 * b_bytes);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int b = tile / p.tiles_per_img, g0 = (tile - b * p.tiles_per_img) * 4;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&empty[s], ph ^ 1u);
        if (elect_one()) {
          uint8_t* st = smem + s * stage_bytes;
          mbar_expect_tx(&full[s], tx);
          tma_load_4d(st, &tmS, &full[s], 0, kc * 32, b, g0);
          tma_load_2d(st + 2 * A_BYTES, &tmHi, &full[s], kc * 32, 0);
          tma_load_2d(st + 2 * A_BYTES + b_bytes, &tmLo, &full[s], kc * 32, 0);
        }
        __syncwarp();
        if (++s == stages) s = 0, ph ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (whole warp loops, one elected lane issues) ----------------
    const uint32_t idesc = umma_idesc(2, 1, 0, 128, p.Cpad);  // tf32, A MN-major, B K-major
    const uint32_t idesc2 = umma_idesc(2, 1, 0, 128, 2 * p.Cpad);
    int s = 0;
    uint32_t ph = 0, local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const uint32_t buf = p.nbuf == 2 ? (local & 1u) : 0u;
      const uint32_t aph = p.nbuf == 2 ? ((local >> 1) & 1u) : (local & 1u);
      ++local;
      mbar_wait(&acce[buf], aph ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem + buf * buf_cols;
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&split[s], ph);
        tc_fence_after();
        if (elect_one()) {
        const uint32_t a_hi = smem_u32(smem + s * stage_bytes), a_lo = a_hi + A_BYTES;
        const uint32_t b_hi = a_hi + 2 * A_BYTES, b_lo = b_hi + (uint32_t)b_bytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // A: MN-major tf32 (SWIZZLE_128B_BASE32B): 32-pixel groups 32*128 B apart, 8 channel rows per MMA in two
          // 4-row halves 512 B apart.  B: K-major SWIZZLE_128B, K advances 32 B inside the 128 B row.
          const uint64_t ah = umma_desc(a_hi + k * 1024, 32 * 128, 512, 1);
          const uint64_t al = umma_desc(a_lo + k * 1024, 32 * 128, 512, 1);
          const uint64_t bh = umma_desc_sw128(b_hi + k * 32, 16, 1024);
          const uint64_t bl = umma_desc_sw128(b_lo + k * 32, 16, 1024);
          const uint32_t acc = (uint32_t)((kc | k) != 0);
          // B_hi and B_lo are stacked along N, so A_hi x [B_hi; B_lo] yields the hi*hi and hi*lo partial sums with ONE
          // instruction (fewer instructions for the same MACs); lo*hi is the second one.
          if (p.nacc == 3) {
            tc_mma<true>(d0, ah, bh, idesc2, acc);                              // columns [0, 2*Cpad): hh | hl
            tc_mma<true>(d0 + 2u * (uint32_t)p.acc_cols, al, bh, idesc, acc);  // columns [2*Cpad, 3*Cpad): lh
          } else {
            // Cpad = 256: three accumulators would need 768 TMEM columns.  The two small terms only ever appear as
            // their sum, so hi*lo and lo*hi share the second accumulator: 2 x 256 = all 512 columns.
            tc_mma<true>(d0, ah, bh, idesc, acc);                               // columns [0, 256): hh
            tc_mma<true>(d0 + (uint32_t)p.acc_cols, ah, bl, idesc, acc);       // columns [256, 512): hl
            tc_mma<true>(d0 + (uint32_t)p.acc_cols, al, bh, idesc, 1u);        //                   += lh
          }
        }
        tc_commit(&empty[s]);
        }
        __syncwarp();
        if (++s == stages) s = 0, ph ^= 1u;
      }
      if (elect_one()) tc_commit(&accf[buf]);
      __syncwarp();
    }
  } else if (warp >= 2 && warp < 6) {
    // ---------------- operand split: A -> (hi in place, lo) ----------------
    const int t = threadIdx.x - 64;  // 0..127: one 128-byte row of each of the 4 pixel groups... (row = group*32 + channel)
    int s = 0;
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < p.kchunks; ++kc) {
        mbar_wait(&full[s], ph);
        uint4* hi = reinterpret_cast<uint4*>(smem + s * stage_bytes + t * 128);
        uint4* lo = reinterpret_cast<uint4*>(smem + s * stage_bytes + A_BYTES + t * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // the split is element-wise, so the swizzle (a permutation of 16/32-byte chunks inside the row) is irrelevant;
          // rotate the chunk order per thread to spread shared-memory banks
          const int jj = (j + t) & 7;
          uint4 v = hi[jj], h, l;
          h.x = v.x & 0xFFFFE000u, h.y = v.y & 0xFFFFE000u, h.z = v.z & 0xFFFFE000u, h.w = v.w & 0xFFFFE000u;
          l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(h.x));
          l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(h.y));
          l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(h.z));
          l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(h.w));
          hi[jj] = h;
          lo[jj] = l;
        }
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        mbar_arrive(&split[s]);
        if (++s == stages) s = 0, ph ^= 1u;
      }
    }
  } else if (warp >= 6) {
    // ---------------- epilogue: sum the three accumulators, scale, first arg-max ----------------
    const int q4 = warp & 3;
    const int row = q4 * 32 + lane;
    uint32_t local = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const uint32_t buf = p.nbuf == 2 ? (local & 1u) : 0u;
      const uint32_t aph = p.nbuf == 2 ? ((local >> 1) & 1u) : (local & 1u);
      ++local;
      mbar_wait(&accf[buf], aph);
      tc_fence_after();
      const uint32_t tbase = tmem + buf * buf_cols + ((uint32_t)(q4 * 32) << 16);
      float best = -INFINITY;
      int besti = 0;
      for (int c0 = 0; c0 < p.Cpad; c0 += 32) {
        uint32_t v[32];
        float f[32];
        tmem_ld32(tbase + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        tmem_ld32(tbase + (uint32_t)(p.acc_cols + c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
        if (p.nacc == 3) {
          tmem_ld32(tbase + (uint32_t)(2 * p.acc_cols + c0), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __uint_as_float(v[j]);
        }
        if (c0 + 32 >= p.Cpad) {  // last TMEM read of the tile
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acce[buf]);
        }
        float ie[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 e4 = __ldg(reinterpret_cast<const float4*>(p.inv_en + c0) + j);
          ie[4 * j] = e4.x, ie[4 * j + 1] = e4.y, ie[4 * j + 2] = e4.z, ie[4 * j + 3] = e4.w;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float val = f[j] * ie[j];
          if (c0 + j < p.C && val > best) best = val, besti = c0 + j;  // strict >: lowest index wins ties
        }
      }
      const int b = tile / p.tiles_per_img;
      const long long pix = (long long)(tile - b * p.tiles_per_img) * 128 + row;
      if (pix < p.hw) p.labels[(long long)b * p.hw + pix] = besti;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem, 512);
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_f32(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                      const cuuint32_t* box, CUtensorMapSwizzle sw) {
  static EncodeTiledFn2 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return set_error(SZN_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    fn = reinterpret_cast<EncodeTiledFn2>(ptr);
  }
  cuuint32_t el[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_b, box, el,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(SZN_ERR_CUDA, "cuTensorMapEncodeTiled failed (szn_embed_argmax)");
  return 0;
}

// returns 1 when the shape is outside the tensor-core path (caller falls back to the CUDA-core kernel), 0 on launch,
// < 0 on error.  scratch: >= 2*Cpad*Dpad + Cpad floats.
int embed_argmax_tc(const float* score, const float* table, int n, int D, long long hw, int C, float* scratch,
                    long long* labels, cudaStream_t st) {
  const int Cpad = C <= 64 ? 64 : C <= 128 ? 128 : 256;
  if (C > 256 || hw % 32 || D < 1 || (reinterpret_cast<uintptr_t>(score) & 15)) return 1;
  const int Dpad = (D + 31) / 32 * 32;
  float* hi = scratch;
  float* lo = hi + (size_t)Cpad * Dpad;
  float* inv_en = lo + (size_t)Cpad * Dpad;
  table_split_kernel<<<Cpad, 128, 0, st>>>(table, C, D, Cpad, Dpad, hi, lo, inv_en);
  if (int e = check_launch("szn_embed_argmax/split")) return e;

  ArgmaxParams p{};
  p.D = D, p.C = C, p.Cpad = Cpad, p.B = n, p.hw = hw;
  p.tiles_per_img = (int)((hw + 127) / 128);
  p.total_tiles = p.tiles_per_img * n;
  p.kchunks = Dpad / 32;
  p.acc_cols = Cpad;
  p.nacc = Cpad == 256 ? 2 : 3;  // Cpad = 256: hi*lo and lo*hi share one accumulator (2 x 256 = 512 TMEM columns)
  p.nbuf = (2 * p.nacc * Cpad <= 512) ? 2 : 1;
  const int stage_bytes = 2 * 128 * 128 + 2 * Cpad * 128;
  int stages = (227 * 1024 - 1024 - 512) / stage_bytes;
  if (stages > 8) stages = 8;
  p.stages = stages;
  p.inv_en = inv_en, p.labels = labels;

  CUtensorMap ts, thi, tlo;
  {
    cuuint64_t d[4] = {32, (cuuint64_t)D, (cuuint64_t)n, (cuuint64_t)(hw / 32)};
    cuuint64_t sb[3] = {(cuuint64_t)hw * 4, (cuuint64_t)D * hw * 4, 128};
    cuuint32_t bx[4] = {32, 32, 1, 4};
    if (int e = encode_f32(&ts, score, 4, d, sb, bx, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return e;
    cuuint64_t d2[2] = {(cuuint64_t)Dpad, (cuuint64_t)Cpad};
    cuuint64_t sb2[1] = {(cuuint64_t)Dpad * 4};
    cuuint32_t bx2[2] = {32, (cuuint32_t)Cpad};
    if (int e = encode_f32(&thi, hi, 2, d2, sb2, bx2, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
    if (int e = encode_f32(&tlo, lo, 2, d2, sb2, bx2, CU_TENSOR_MAP_SWIZZLE_128B)) return e;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(embed_argmax_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return set_error(SZN_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)(p.total_tiles < sms ? p.total_tiles : sms);
  const size_t smem = (size_t)stages * stage_bytes + 1024 + 512;
  embed_argmax_tc_kernel<<<grid, 320, smem, st>>>(ts, thi, tlo, p);
  return check_launch("szn_embed_argmax/tc");
}

}  // namespace szn
