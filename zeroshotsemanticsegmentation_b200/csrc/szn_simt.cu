// CUDA-core kernels of the trunk that are HBM-bound or too thin for tensor tiles:
// conv1_1 (Cin=3, pad=100; models.py:43), 2x2 ceil-mode max-pool fwd/bwd (models.py:47..81),
// bias gradients, weight layout packing, Dropout2d channel masks (models.py:86,91).
#include "szn_internal.h"
#include <stdlib.h>
#include "szn_ptx.cuh"
#include "szn_store.cuh"

namespace szn {

static char g_err[512] = "";
static long long g_launches = 0;
int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof g_err, "%s", msg);
  return code;
}
void count_launch() { ++g_launches; }

template <typename T>
__device__ __forceinline__ T from_float(float f);
template <>
__device__ __forceinline__ float from_float<float>(float f) { return to_tf32(f); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float f) { return __float2bfloat16_rn(f); }
template <typename T>
__device__ __forceinline__ float as_float(T v);
template <>
__device__ __forceinline__ float as_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ float as_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------------------------------------
// conv1_1 forward: x NCHW fp32 [B,3,H,W] -> y NHWC [B,Ho,Wo,64], Ho = H + 2*pad - 2, bias + ReLU.
// One thread = one output pixel x 64 channels; the CTA's 128 pixels are contiguous in NHWC, so the
// tile is staged in shared memory and written back with coalesced 16-byte stores.
// Pixels whose 3x3 window lies entirely in the padding (47% of them at 512x512) are relu(bias).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) conv1_1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w /*OIHW [64][3][3][3]*/,
                                                          const float* __restrict__ bias, void* __restrict__ y, int B, int H,
                                                          int W, int Ho, int Wo, int pad) {
  using S = Store<T>;
  constexpr int VN = S::VN;                          // channels per vector
  constexpr int ROWB = (int)S::row_bytes(64);        // bytes of one pixel row (64 channels)
  constexpr int NCH = ROWB / 16;                     // 16-byte chunks per row: 16 (fp32, split) / 8 (bf16)
  constexpr int NV = 64 / VN;                        // channel vectors per row
  __shared__ __align__(16) float sw_[27 * 64];  // [k][co]
  __shared__ float sb[64];
  __shared__ __align__(16) uint4 tile[128 * NCH];
  for (int i = threadIdx.x; i < 27 * 64; i += 128) {
    const int co = i & 63, k = i >> 6;
    const int tap = k / 3, c = k - tap * 3;
    sw_[k * 64 + co] = w[co * 27 + c * 9 + tap];  // OIHW -> [tap][ci] order
  }
  if (threadIdx.x < 64) sb[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const long long total = (long long)B * Ho * Wo;
  const long long p0 = (long long)blockIdx.x * 128;
  const long long pix = p0 + threadIdx.x;
  if (pix < total) {
    const int xo = (int)(pix % Wo);
    const int yo = (int)((pix / Wo) % Ho);
    const int b = (int)(pix / ((long long)Wo * Ho));
    float in[27];
    bool any = false;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int yi = yo + r - pad, xi = xo + s - pad;
        const bool ok = yi >= 0 && yi < H && xi >= 0 && xi < W;
        any |= ok;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          in[(r * 3 + s) * 3 + c] = ok ? __ldg(x + (((long long)b * 3 + c) * H + yi) * W + xi) : 0.f;
      }
    // the thread's row goes to row threadIdx.x of the tile as 16-byte chunks, chunk c stored at c ^ (row & (NCH-1)):
    // a plain [row][channel] layout makes all 32 lanes hit the same bank (row stride = a multiple of 128 bytes)
    uint4* trow = tile + threadIdx.x * NCH;
    const int sw = threadIdx.x & (NCH - 1);
#pragma unroll 1
    for (int cv = 0; cv < NV; ++cv) {
      float a[VN];
#pragma unroll
      for (int q = 0; q < VN; q += 4) {
        const int c4 = cv * VN + q;
        float a0 = sb[c4], a1 = sb[c4 + 1], a2 = sb[c4 + 2], a3 = sb[c4 + 3];
        if (any) {
#pragma unroll
          for (int k = 0; k < 27; ++k) {
            const float4 wv = *reinterpret_cast<const float4*>(sw_ + k * 64 + c4);
            a0 = fmaf(in[k], wv.x, a0);
            a1 = fmaf(in[k], wv.y, a1);
            a2 = fmaf(in[k], wv.z, a2);
            a3 = fmaf(in[k], wv.w, a3);
          }
        }
        a[q] = fmaxf(a0, 0.f), a[q + 1] = fmaxf(a1, 0.f), a[q + 2] = fmaxf(a2, 0.f), a[q + 3] = fmaxf(a3, 0.f);
      }
      const typename S::Raw r = S::from_float(a);
      trow[cv ^ sw] = r.a;
      if constexpr (sizeof(typename S::Raw) == 32) trow[(cv + NV) ^ sw] = r.b;  // lo plane: chunks NV .. 2 NV - 1
    }
  }
  __syncthreads();
  long long npix = total - p0;
  if (npix > 128) npix = 128;
  const int n16 = (int)(npix * NCH);
  uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(y) + (size_t)p0 * ROWB);
  for (int i = threadIdx.x; i < n16; i += 128) {
    const int row = i / NCH, c = i - row * NCH;
    __stcs(dst + i, tile[row * NCH + (c ^ (row & (NCH - 1)))]);
  }
}

// conv1_1 weight gradient: dw[64][27] += sum_pixels dy[p][co] * x[p + tap - pad][ci]   (fp32 atomics)
// Only output pixels whose 3x3 window touches the image contribute (x is zero elsewhere): with pad = 100 that is the
// central (H+2) x (W+2) window, 52 % of the 710^2 map at 512^2.
// Work item = 64 consecutive output pixels of one row.  576 threads = 64 output channels x 9 (input channel, filter row)
// pairs; a thread slides a 3-wide window over the staged input row, so a pixel costs it one dY load (coalesced over the
// channels), one shared-memory load and 3 FMAs for its three filter columns.  CTAs are persistent and keep their 3
// partial sums in registers across all work items: 1728 atomics per CTA in total.
constexpr int C11_SEG = 64;
// The split format is staged as fp32 (hi + lo added on the way into shared memory).
template <typename T>
struct C11Stage {
  typedef T type;
};
template <>
struct C11Stage<SplitBf16> {
  typedef float type;
};
template <typename T>
__global__ void __launch_bounds__(576) conv1_1_wgrad_kernel(const float* __restrict__ x, const void* __restrict__ dy_,
                                                            float* __restrict__ dw, int B, int H, int W, int Ho, int Wo,
                                                            int pad, int ylo, int xlo, int wy, int nseg) {
  constexpr bool SPLIT = sizeof(typename Store<T>::Raw) == 32;
  typedef typename C11Stage<T>::type ST;             // element type of the staged dY tile
  constexpr int VN = Store<T>::VN;                   // dY channels per (plane) 16-byte load
  constexpr int NV = C11_SEG * 64 / VN;              // channel vectors in one dY tile (64 px x 64 channels)
  constexpr int VPT = (NV + 575) / 576;              // vectors per thread
  constexpr int XE = 9 * (C11_SEG + 2), XPT = (XE + 575) / 576;
  __shared__ __align__(16) ST sdy[2][C11_SEG * 64];
  const uint8_t* dy = reinterpret_cast<const uint8_t*>(dy_);
  __shared__ float xs[2][9][C11_SEG + 2];  // [buffer][ci*3 + r][x]: the three input rows under this output row
  const int co = threadIdx.x & 63, cr = threadIdx.x >> 6;  // cr = ci*3 + r
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  const long long items = (long long)B * wy * nseg;
  typename Store<T>::Raw rdy[VPT];
  float rx[XPT];
  // software pipeline: the global loads of item i+1 are in flight while item i is reduced out of shared memory
  auto fetch = [&](long long it) {
    const int seg = (int)(it % nseg);
    const int yo = ylo + (int)((it / nseg) % wy);
    const int b = (int)(it / ((long long)nseg * wy));
    const int xo0 = xlo + seg * C11_SEG;
    int npx = Wo - xo0;
    if (npx > C11_SEG) npx = C11_SEG;
    const long long row0 = ((long long)b * Ho + yo) * Wo + xo0;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * 576;  // vector i = pixel i / (64 / VN), channel vector i % (64 / VN)
      const int px = i / (64 / VN), cv = i - px * (64 / VN);
      rdy[v] = (px < npx) ? Store<T>::template load_raw<true>(row_ptr<T>(dy, row0 + px, 64), 64, cv)
                          : Store<T>::zero();  // pixels past the row end count as 0
    }
#pragma unroll
    for (int v = 0; v < XPT; ++v) {
      const int i = threadIdx.x + v * 576;
      float val = 0.f;
      if (i < XE) {
        const int row = i / (C11_SEG + 2), col = i - row * (C11_SEG + 2);
        const int ci = row / 3, r = row - ci * 3;
        const int yi = yo + r - pad, xi = xo0 - pad + col;
        if (yi >= 0 && yi < H && xi >= 0 && xi < W) val = __ldg(x + (((long long)b * 3 + ci) * H + yi) * W + xi);
      }
      rx[v] = val;
    }
  };
  int buf = 0;
  if ((long long)blockIdx.x < items) fetch(blockIdx.x);
  for (long long it = blockIdx.x; it < items; it += gridDim.x, buf ^= 1) {
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * 576;
      if (i < NV) {
        if constexpr (SPLIT) {
          float f[VN];
          Store<T>::to_float(rdy[v], f);
          float4* d4 = reinterpret_cast<float4*>(sdy[buf] + i * VN);
          d4[0] = make_float4(f[0], f[1], f[2], f[3]);
          d4[1] = make_float4(f[4], f[5], f[6], f[7]);
        } else {
          reinterpret_cast<uint4*>(sdy[buf])[i] = rdy[v].a;
        }
      }
    }
#pragma unroll
    for (int v = 0; v < XPT; ++v) {
      const int i = threadIdx.x + v * 576;
      if (i < XE) (&xs[buf][0][0])[i] = rx[v];
    }
    __syncthreads();  // double-buffered: nobody still reads this buffer (its previous readers passed the last barrier)
    if (it + gridDim.x < items) fetch(it + gridDim.x);
    const ST* d = sdy[buf] + co;
    float w0 = xs[buf][cr][0], w1 = xs[buf][cr][1];
#pragma unroll 16
    for (int px = 0; px < C11_SEG; ++px) {
      const float dv = as_float<ST>(d[px * 64]);
      const float w2 = xs[buf][cr][px + 2];
      acc0 = fmaf(dv, w0, acc0);
      acc1 = fmaf(dv, w1, acc1);
      acc2 = fmaf(dv, w2, acc2);
      w0 = w1, w1 = w2;
    }
  }
  const int ci = cr / 3, r = cr - ci * 3;
  float* o = dw + co * 27 + ci * 9 + r * 3;  // OIHW
  atomicAdd(o, acc0);
  atomicAdd(o + 1, acc1);
  atomicAdd(o + 2, acc2);
}

// Re-blocked form (the default, SZN_CONV1_1_WGRAD_V2): 192 threads = 64 output channels x 3 input channels, and a thread keeps all NINE taps
// of its (co, ci) pair in registers.  A dY value is read from shared memory once per input channel instead of once per
// (input channel, filter row) pair, and the three input rows come in as broadcast 16-byte loads of four pixels: 7 LDS per
// 36 FMAs where the kernel above needs 24, which moves it from the shared-memory pipe to the FMA pipe.  Only
// ceil(npx / 4) pixel quads of a segment are visited, so the 2-pixel tail segment of a 514-pixel row costs 1/16 of a
// full one.
constexpr int C11_NT = 192;
constexpr int C11_XROW = C11_SEG + 4;  // staged input row: 64 + 2 halo columns, padded to whole float4s
template <typename T>
__global__ void __launch_bounds__(C11_NT) conv1_1_wgrad_v2_kernel(const float* __restrict__ x, const void* __restrict__ dy_,
                                                                 float* __restrict__ dw, int B, int H, int W, int Ho, int Wo,
                                                                 int pad, int ylo, int xlo, int wy, int nseg) {
  constexpr bool SPLIT = sizeof(typename Store<T>::Raw) == 32;
  typedef typename C11Stage<T>::type ST;
  constexpr int VN = Store<T>::VN;
  constexpr int NV = C11_SEG * 64 / VN;
  constexpr int VPT = (NV + C11_NT - 1) / C11_NT;
  constexpr int XE = 9 * C11_XROW, XPT = (XE + C11_NT - 1) / C11_NT;
  __shared__ __align__(16) ST sdy[2][C11_SEG * 64];
  __shared__ __align__(16) float xs[2][9][C11_XROW];  // [buffer][ci*3 + r][x]
  const uint8_t* dy = reinterpret_cast<const uint8_t*>(dy_);
  const int co = threadIdx.x & 63, ci = threadIdx.x >> 6;
  float acc[3][3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int t = 0; t < 3; ++t) acc[r][t] = 0.f;
  const long long items = (long long)B * wy * nseg;
  typename Store<T>::Raw rdy[VPT];
  float rx[XPT];
  auto seg_px = [&](long long it) {  // valid pixels of work item `it`
    const int n = Wo - (xlo + (int)(it % nseg) * C11_SEG);
    return n > C11_SEG ? C11_SEG : n;
  };
  auto fetch = [&](long long it) {
    const int seg = (int)(it % nseg);
    const int yo = ylo + (int)((it / nseg) % wy);
    const int b = (int)(it / ((long long)nseg * wy));
    const int xo0 = xlo + seg * C11_SEG;
    const int npx = seg_px(it);
    const long long row0 = ((long long)b * Ho + yo) * Wo + xo0;
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * C11_NT;
      const int px = i / (64 / VN), cv = i - px * (64 / VN);
      rdy[v] = (i < NV && px < npx) ? Store<T>::template load_raw<true>(row_ptr<T>(dy, row0 + px, 64), 64, cv)
                                    : Store<T>::zero();  // pixels past the row end count as 0
    }
#pragma unroll
    for (int v = 0; v < XPT; ++v) {
      const int i = threadIdx.x + v * C11_NT;
      float val = 0.f;
      if (i < XE) {
        const int row = i / C11_XROW, col = i - row * C11_XROW;
        const int c = row / 3, r = row - c * 3;
        const int yi = yo + r - pad, xi = xo0 - pad + col;
        if (yi >= 0 && yi < H && xi >= 0 && xi < W) val = __ldg(x + (((long long)b * 3 + c) * H + yi) * W + xi);
      }
      rx[v] = val;
    }
  };
  int buf = 0;
  if ((long long)blockIdx.x < items) fetch(blockIdx.x);
  for (long long it = blockIdx.x; it < items; it += gridDim.x, buf ^= 1) {
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
      const int i = threadIdx.x + v * C11_NT;
      if (i < NV) {
        if constexpr (SPLIT) {
          float f[VN];
          Store<T>::to_float(rdy[v], f);
          float4* d4 = reinterpret_cast<float4*>(sdy[buf] + i * VN);
          d4[0] = make_float4(f[0], f[1], f[2], f[3]);
          d4[1] = make_float4(f[4], f[5], f[6], f[7]);
        } else {
          reinterpret_cast<uint4*>(sdy[buf])[i] = rdy[v].a;
        }
      }
    }
#pragma unroll
    for (int v = 0; v < XPT; ++v) {
      const int i = threadIdx.x + v * C11_NT;
      if (i < XE) (&xs[buf][0][0])[i] = rx[v];
    }
    __syncthreads();  // double-buffered: the previous readers of this buffer passed the last barrier
    if (it + gridDim.x < items) fetch(it + gridDim.x);
    const int nq = (seg_px(it) + 3) >> 2;  // pixel quads with at least one valid pixel (the staged tail is zero)
    const ST* d = sdy[buf] + co;
    const float4* x0 = reinterpret_cast<const float4*>(xs[buf][ci * 3]);
    const float4* x1 = reinterpret_cast<const float4*>(xs[buf][ci * 3 + 1]);
    const float4* x2 = reinterpret_cast<const float4*>(xs[buf][ci * 3 + 2]);
    float4 cur[3] = {x0[0], x1[0], x2[0]};
#pragma unroll 2
    for (int k = 0; k < nq; ++k) {
      const float4 nxt[3] = {x0[k + 1], x1[k + 1], x2[k + 1]};
      float dv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) dv[j] = as_float<ST>(d[(4 * k + j) * 64]);
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float wv[6] = {cur[r].x, cur[r].y, cur[r].z, cur[r].w, nxt[r].x, nxt[r].y};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int t = 0; t < 3; ++t) acc[r][t] = fmaf(dv[j], wv[j + t], acc[r][t]);
        cur[r] = nxt[r];
      }
    }
  }
  float* o = dw + co * 27 + ci * 9;  // OIHW
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int t = 0; t < 3; ++t) atomicAdd(o + r * 3 + t, acc[r][t]);
}

// ------------------------------------------------------------------------------------------------
// MaxPool2d(2, stride 2, ceil_mode=True) on NHWC, one channel vector (Store<T>::VN channels) per thread
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pool_fwd_kernel(const void* __restrict__ in, void* __restrict__ out, int B, int H, int W, int C, int Ho, int Wo) {
  using S = Store<T>;
  constexpr int VN = S::VN;
  const int cv = C / VN;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long r = i / cv;
    const int xo = (int)(r % Wo);
    r /= Wo;
    const int yo = (int)(r % Ho);
    const int b = (int)(r / Ho);
    // four independent loads in flight per thread; window positions outside the map (ceil mode) are clamped onto the last
    // row / column, which repeats a value of the window and leaves the maximum unchanged
    typename S::Raw raw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int y = 2 * yo + (q >> 1), x = 2 * xo + (q & 1);
      y = y < H ? y : H - 1, x = x < W ? x : W - 1;
      raw[q] = S::template load_raw<false>(row_ptr<T>(in, ((long long)b * H + y) * W + x, C), C, c);
    }
    float m[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[VN];
      S::to_float(raw[q], v);
#pragma unroll
      for (int j = 0; j < VN; ++j) m[j] = fmaxf(m[j], v[j]);
    }
    // tf32 / bf16: the maximum is one of the stored values, re-encoding it is exact.  split: hi + lo of the winner is
    // re-split, which reproduces the same fp32 value to 2^-17 (the planes themselves may differ in the last bit).
    S::template store_raw<false>(row_ptr<T>(out, ((long long)b * Ho + yo) * Wo + xo, C), C, c, S::from_float(m));
  }
}

// dY[b,y,x,c] = dP[b,y/2,x/2,c] if Y[b,y,x,c] is the FIRST maximum of its window (scan order, like ATen) and > 0 (ReLU).
// One thread = one 2x2 window x one channel vector: every byte of Y, dP and dY moves exactly once.
template <typename T>
__global__ void __launch_bounds__(256) pool_bwd_kernel(const void* __restrict__ yin, const void* __restrict__ dp,
                                                       void* __restrict__ dy, int B, int H, int W, int C, int Ho, int Wo,
                                                       int relu_gate, float* __restrict__ col_sum) {
  using S = Store<T>;
  constexpr int VN = S::VN;
  const int cv = C / VN;
  // per-channel sums of dy (= of the routed dp values that pass the gate): the producer conv's bias gradient.
  // 256 % cv == 0 is guaranteed by the host when col_sum is given, so a thread keeps ONE channel vector for the whole
  // grid-stride loop and accumulates in registers; one shared-memory pass and C atomics per block at the end.
  float csum[VN];
#pragma unroll
  for (int j = 0; j < VN; ++j) csum[j] = 0.f;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long r = i / cv;
    const int xo = (int)(r % Wo);
    r /= Wo;
    const int yo = (int)(r % Ho);
    const int b = (int)(r / Ho);
    float yv[4][VN];
    bool have[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = 2 * yo + (q >> 1), xx = 2 * xo + (q & 1);
      have[q] = yy < H && xx < W;
      if (have[q]) {
        S::to_float(S::template load_raw<true>(row_ptr<T>(yin, ((long long)b * H + yy) * W + xx, C), C, c), yv[q]);
      } else {
#pragma unroll
        for (int j = 0; j < VN; ++j) yv[q][j] = 0.f;
      }
    }
    float g[VN];
    S::to_float(S::template load_raw<true>(row_ptr<T>(dp, ((long long)b * Ho + yo) * Wo + xo, C), C, c), g);
    float o[4][VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      float best = -INFINITY;
      int win = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (have[q] && yv[q][j] > best) best = yv[q][j], win = q;  // strict >: the first maximum in scan order keeps the gradient
      const bool pass = !relu_gate || best > 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][j] = (pass && q == win) ? g[j] : 0.f;
      if (pass) csum[j] += g[j];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = 2 * yo + (q >> 1), xx = 2 * xo + (q & 1);
      // g is a stored value: re-encoding it is exact (tf32 / bf16) or the same fp32 value to 2^-17 (split)
      if (have[q]) S::template store_raw<true>(row_ptr<T>(dy, ((long long)b * H + yy) * W + xx, C), C, c, S::from_float(o[q]));
    }
  }
  if (col_sum) {
    __shared__ float red[256 * VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) red[threadIdx.x * VN + j] = csum[j];
    __syncthreads();
    // threads t, t + cv, t + 2cv, ... hold the same channel vector (gridDim.x * 256 is a multiple of cv)
    if ((int)threadIdx.x < cv) {
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        float t = 0.f;
        for (int k = threadIdx.x; k < 256; k += cv) t += red[k * VN + j];
        atomicAdd(col_sum + (((blockIdx.x * 256ll) + threadIdx.x) % cv) * VN + j, t);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The same pool pair with a one-byte ROUTING CODE per pooled element instead of a second read of the pre-pool tensor:
// code = winner (0..3, scan order, first maximum like ATen) | 4 if the maximum is > 0 (the producer's ReLU gate).
// The forward pass writes it next to the pooled value (+1/16 of the pre-pool bytes in fp32 storage); the backward pass
// reads dP and the code and writes dY: 1.31x the pre-pool tensor instead of 2.25x, and the pre-pool activation is not kept
// for the backward pass at all.  Winner and gate are decided by the comparisons of pool_bwd_kernel, so dY is bit-identical.
// ------------------------------------------------------------------------------------------------
template <int VN>
struct CodeVec;
template <>
struct CodeVec<4> {
  typedef uint32_t type;
};
template <>
struct CodeVec<8> {
  typedef uint2 type;
};

template <typename T>
__global__ void __launch_bounds__(256) pool_fwd_code_kernel(const void* __restrict__ in, void* __restrict__ out,
                                                            uint8_t* __restrict__ code, int B, int H, int W, int C, int Ho,
                                                            int Wo) {
  using S = Store<T>;
  constexpr int VN = S::VN;
  const int cv = C / VN;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long r = i / cv;
    const int xo = (int)(r % Wo);
    r /= Wo;
    const int yo = (int)(r % Ho);
    const int b = (int)(r / Ho);
    // All four loads are issued before the first comparison (four independent requests in flight per thread instead of a
    // load -> compare chain behind four branches).  Window positions outside the map (ceil mode, odd H or W) are clamped
    // onto the last row / column: that repeats a value seen EARLIER in scan order, which a strict > never lets win.
    typename S::Raw raw[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      int y = 2 * yo + (q >> 1), x = 2 * xo + (q & 1);
      y = y < H ? y : H - 1, x = x < W ? x : W - 1;
      raw[q] = S::template load_raw<true>(row_ptr<T>(in, ((long long)b * H + y) * W + x, C), C, c);
    }
    float best[VN];
    uint32_t win[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) best[j] = -INFINITY, win[j] = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[VN];
      S::to_float(raw[q], v);
#pragma unroll
      for (int j = 0; j < VN; ++j)
        if (v[j] > best[j]) best[j] = v[j], win[j] = q;  // strict >: the first maximum in scan order wins
    }
    const long long orow = ((long long)b * Ho + yo) * Wo + xo;
    S::template store_raw<false>(row_ptr<T>(out, orow, C), C, c, S::from_float(best));
    uint32_t w[VN / 4];
#pragma unroll
    for (int k = 0; k < VN / 4; ++k) {
      w[k] = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) w[k] |= (win[4 * k + j] | (best[4 * k + j] > 0.f ? 4u : 0u)) << (8 * j);
    }
    typename CodeVec<VN>::type* cp = reinterpret_cast<typename CodeVec<VN>::type*>(code + orow * C) + c;
    if constexpr (VN == 4) *cp = w[0];
    else *cp = make_uint2(w[0], w[1]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pool_bwd_code_kernel(const uint8_t* __restrict__ code, const void* __restrict__ dp,
                                                            void* __restrict__ dy, int B, int H, int W, int C, int Ho, int Wo,
                                                            int relu_gate, float* __restrict__ col_sum) {
  using S = Store<T>;
  constexpr int VN = S::VN;
  const int cv = C / VN;
  float csum[VN];  // see pool_bwd_kernel: a thread keeps one channel vector for the whole grid-stride loop
#pragma unroll
  for (int j = 0; j < VN; ++j) csum[j] = 0.f;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long r = i / cv;
    const int xo = (int)(r % Wo);
    r /= Wo;
    const int yo = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const long long orow = ((long long)b * Ho + yo) * Wo + xo;
    float g[VN];
    S::to_float(S::template load_raw<true>(row_ptr<T>(dp, orow, C), C, c), g);
    uint32_t w[VN / 4];
    const typename CodeVec<VN>::type* cp = reinterpret_cast<const typename CodeVec<VN>::type*>(code + orow * C) + c;
    if constexpr (VN == 4) {
      w[0] = __ldcs(cp);
    } else {
      const uint2 t = __ldcs(cp);
      w[0] = t.x, w[1] = t.y;
    }
    float o[4][VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) {
      const uint32_t cd = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
      const bool pass = !relu_gate || (cd & 4u);
#pragma unroll
      for (int q = 0; q < 4; ++q) o[q][j] = (pass && (cd & 3u) == (uint32_t)q) ? g[j] : 0.f;
      if (pass) csum[j] += g[j];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = 2 * yo + (q >> 1), xx = 2 * xo + (q & 1);
      if (yy < H && xx < W) S::template store_raw<true>(row_ptr<T>(dy, ((long long)b * H + yy) * W + xx, C), C, c, S::from_float(o[q]));
    }
  }
  if (col_sum) {
    __shared__ float red[256 * VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) red[threadIdx.x * VN + j] = csum[j];
    __syncthreads();
    if ((int)threadIdx.x < cv) {
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        float t = 0.f;
        for (int k = threadIdx.x; k < 256; k += cv) t += red[k * VN + j];
        atomicAdd(col_sum + (((blockIdx.x * 256ll) + threadIdx.x) % cv) * VN + j, t);
      }
    }
  }
}

// db[c] += sum over rows of dy[row][c]   (dy row stride ld).  HBM-bound single pass: a thread owns one channel vector
// and walks rows with a block-wide stride, 4 loads in flight; partial sums meet in shared memory, one atomic per
// (block, channel).
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(const void* __restrict__ dy, float* __restrict__ db, long long rows,
                                                        int C, long long ld, long long rows_per_block) {
  using S = Store<T>;
  constexpr int VN = S::VN;
  __shared__ float red[256 * VN];
  const int cv = (C + VN - 1) / VN;          // channel vectors per row (C % VN == 0 is checked by the host)
  const int lanes = cv < 256 ? cv : 256;      // threads along the channel axis
  const int rstep = 256 / lanes;              // rows handled concurrently by the block
  const int tc = threadIdx.x % lanes, tr = threadIdx.x / lanes;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  for (int c0 = 0; c0 < cv; c0 += lanes) {
    const int c = c0 + tc;
    float acc[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) acc[j] = 0.f;
    if (c < cv && tr < rstep) {
#pragma unroll 4
      for (long long r = r0 + tr; r < r1; r += rstep) {
        float v[VN];
        S::to_float(S::template load_raw<true>(row_ptr<T>(dy, r, ld), (int)ld, c), v);
#pragma unroll
        for (int j = 0; j < VN; ++j) acc[j] += v[j];
      }
    }
#pragma unroll
    for (int j = 0; j < VN; ++j) red[threadIdx.x * VN + j] = acc[j];
    __syncthreads();
    if (tr == 0 && c < cv) {
#pragma unroll
      for (int j = 0; j < VN; ++j) {
        float t = 0.f;
        for (int k = 0; k < rstep; ++k) t += red[(k * lanes + tc) * VN + j];
        if (c * VN + j < C) atomicAdd(db + c * VN + j, t);
      }
    }
    __syncthreads();
  }
}

// one packed weight element: tf32-rounded fp32, bf16, or (split) bf16 hi at out[i] and lo at out[plane + i]
template <typename T>
__device__ __forceinline__ void put_weight(void* out, long long i, long long plane, float v) {
  if constexpr (sizeof(typename Store<T>::Raw) == 32) {
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    reinterpret_cast<__nv_bfloat16*>(out)[i] = hi;
    reinterpret_cast<__nv_bfloat16*>(out)[plane + i] = lo;
  } else {
    reinterpret_cast<T*>(out)[i] = from_float<T>(v);
  }
}

// OIHW fp32 -> [O_pad][R*S][I] T   (rows >= O are zero; fp32 values are rounded to tf32; split: two such planes, hi then lo)
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, void* __restrict__ out, int O, int I, int RS, int O_pad) {
  const long long total = (long long)O_pad * RS * I;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % I);
    const int rs = (int)((i / I) % RS);
    const int o = (int)(i / ((long long)I * RS));
    put_weight<T>(out, i, total, o < O ? w[((long long)o * I + ci) * RS + rs] : 0.f);
  }
}
// OIHW fp32 -> the transposed layouts the data-gradient GEMMs read (K = output channel, contiguous):
//   mode 0: out[ci][(R-1-r)*S + (S-1-s)][co]   dgrad as a forward conv of dY with flipped taps
//   mode 1: out[(r*S+s)*I + ci][co]            dgrad as one GEMM producing per-tap columns (then szn_col2im)
// co >= O is zero-filled up to O_pad.
template <typename T>
__global__ void pack_weight_dgrad_kernel(const float* __restrict__ w, void* __restrict__ out, int O, int I, int R, int S,
                                         int O_pad, int mode) {
  const int RS = R * S;
  const long long total = (long long)I * RS * O_pad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % O_pad);
    const long long row = i / O_pad;
    int ci, tap;
    if (mode == 0) {
      ci = (int)(row / RS);
      tap = RS - 1 - (int)(row - (long long)ci * RS);  // flipped in both r and s
    } else {
      tap = (int)(row / I);
      ci = (int)(row - (long long)tap * I);
    }
    put_weight<T>(out, i, total, co < O ? w[((long long)co * I + ci) * RS + tap] : 0.f);
  }
}

// dx[b,Y,X,ci] = sum_{r,s} dcol[b, Y-r, X-s, (r*S+s)*C + ci]   (valid conv, pad 0): the scatter-free transpose of im2col
template <typename T>
__global__ void col2im_kernel(const void* __restrict__ dcol, void* __restrict__ dx, int B, int H, int W, int C, int R, int S) {
  using St = Store<T>;
  constexpr int VN = St::VN;
  const int Ho = H - R + 1, Wo = W - S + 1, cv = C / VN;
  const long long total = (long long)B * H * W * cv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long q = i / cv;
    const int X = (int)(q % W);
    q /= W;
    const int Y = (int)(q % H);
    const int b = (int)(q / H);
    float acc[VN];
#pragma unroll
    for (int j = 0; j < VN; ++j) acc[j] = 0.f;
    for (int r = 0; r < R; ++r) {
      const int y = Y - r;
      if (y < 0 || y >= Ho) continue;
      for (int sx = 0; sx < S; ++sx) {
        const int x = X - sx;
        if (x < 0 || x >= Wo) continue;
        // a dcol pixel row holds R*S*C channels; this tap's C channels start at channel vector (r*S+sx)*cv
        float v[VN];
        St::to_float(St::template load_raw<false>(row_ptr<T>(dcol, ((long long)b * Ho + y) * Wo + x, (long long)R * S * C),
                                                  R * S * C, (r * S + sx) * cv + c), v);
#pragma unroll
        for (int j = 0; j < VN; ++j) acc[j] += v[j];
      }
    }
    St::template store_raw<false>(row_ptr<T>(dx, ((long long)b * H + Y) * W + X, C), C, c, St::from_float(acc));
  }
}

// [O_pad][R*S][I] fp32 -> OIHW fp32
__global__ void unpack_wgrad_kernel(const float* __restrict__ dw, float* __restrict__ g, int O, int I, int RS) {
  const long long total = (long long)O * RS * I;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int rs = (int)(i % RS);
    const int ci = (int)((i / RS) % I);
    const int o = (int)(i / ((long long)I * RS));
    g[i] = dw[((long long)o * RS + rs) * I + ci];
  }
}

// Dropout2d(p=0.5) channel multipliers: scale[b][c] in {0, 2}  (counter-based hash RNG)
__global__ void dropout_scale_kernel(float* __restrict__ scale, int n, unsigned long long seed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  scale[i] = (z >> 40) & 1ull ? 2.f : 0.f;
}

// fp32 [rows][C] -> T [rows][C] (used for casting small tensors; the split format needs the row length)
template <typename T>
__global__ void cast_kernel(const float* __restrict__ in, void* __restrict__ out, long long n, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    Store<T>::store_elem(out, i / C, C, (int)(i % C), in[i]);
}

// torch.optim.SGD step (train.py:126-129: momentum .99, weight_decay 5e-4, biases lr x2 / no decay), one fused pass:
//   d = g + wd * p;  buf = first ? d : momentum * buf + d;  p -= lr * buf        (dampening 0, no Nesterov)
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                  long long n, float lr, float momentum, float wd, int first) {
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mv = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(m)[i];
    float d;
    d = fmaf(wd, pv.x, gv.x), mv.x = first ? d : fmaf(momentum, mv.x, d), pv.x = fmaf(-lr, mv.x, pv.x);
    d = fmaf(wd, pv.y, gv.y), mv.y = first ? d : fmaf(momentum, mv.y, d), pv.y = fmaf(-lr, mv.y, pv.y);
    d = fmaf(wd, pv.z, gv.z), mv.z = first ? d : fmaf(momentum, mv.z, d), pv.z = fmaf(-lr, mv.z, pv.z);
    d = fmaf(wd, pv.w, gv.w), mv.w = first ? d : fmaf(momentum, mv.w, d), pv.w = fmaf(-lr, mv.w, pv.w);
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(p)[i] = pv;
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = fmaf(wd, p[i], g[i]);
    const float b = first ? d : fmaf(momentum, m[i], d);
    m[i] = b;
    p[i] = fmaf(-lr, b, p[i]);
  }
}

// Adam (torch.optim.Adam, amsgrad off): one pass instead of torch's six elementwise passes.  step_size = lr / (1 - b1^t),
// inv_bc2_sqrt is passed as bc2_sqrt = sqrt(1 - b2^t) and applied as a division, like torch's single-tensor path.
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float b1, float b2, float one_minus_b1,
                                            float one_minus_b2, float eps, float wd, float step_size, float bc2_sqrt) {
  g = fmaf(wd, p, g);
  m = fmaf(one_minus_b1, g - m, m);            // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(one_minus_b2 * g, g, b2 * v);       // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = fmaf(-step_size, m / denom, p);
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float b1, float b2, float eps,
                                                   float wd, float step_size, float bc2_sqrt) {
  const float omb1 = 1.f - b1, omb2 = 1.f - b2;
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_update(pv.x, gv.x, mv.x, vv.x, b1, b2, omb1, omb2, eps, wd, step_size, bc2_sqrt);
    adam_update(pv.y, gv.y, mv.y, vv.y, b1, b2, omb1, omb2, eps, wd, step_size, bc2_sqrt);
    adam_update(pv.z, gv.z, mv.z, vv.z, b1, b2, omb1, omb2, eps, wd, step_size, bc2_sqrt);
    adam_update(pv.w, gv.w, mv.w, vv.w, b1, b2, omb1, omb2, eps, wd, step_size, bc2_sqrt);
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(p)[i] = pv;
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    adam_update(p[i], g[i], m[i], v[i], b1, b2, omb1, omb2, eps, wd, step_size, bc2_sqrt);
}

static int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace szn
using namespace szn;

extern "C" const char* szn_last_error(void) { return g_err; }
extern "C" long long szn_launch_count(void) { return g_launches; }
extern "C" int szn_abi_version(void) { return 2; }

#define DISPATCH_T(dtype, CALL)                 \
  do {                                          \
    if ((dtype) == SZN_BF16) {                  \
      typedef __nv_bfloat16 T;                  \
      CALL;                                     \
    } else if ((dtype) == SZN_F32) {            \
      typedef float T;                          \
      CALL;                                     \
    } else if ((dtype) == SZN_F32X3) {          \
      typedef SplitBf16 T;                      \
      CALL;                                     \
    } else {                                    \
      return set_error(SZN_ERR_ARG, "bad dtype"); \
    }                                           \
  } while (0)

namespace szn {
int conv1_1_fwd_tc(int dtype, const float* x, const float* w, const float* bias, void* y, int B, int H, int W, int pad,
                   cudaStream_t st);  // szn_conv1_1_tc.cu
}

extern "C" int szn_conv1_1_fwd(int dtype, const float* x, const float* w_oihw, const float* bias, void* y, int B, int H,
                               int W, int pad, void* stream) {
  // TF32 / bf16 storage: the tensor-core kernel (3 x TF32 GEMM over im2col rows built in shared memory: 0.50 -> 0.42 ms at
  // B = 8, 512 x 512).  The fp32-grade mode keeps the CUDA-core kernel below, whose fp32 FMAs are exact to the last bit:
  // with 3 x TF32 products (2^-21) in the FIRST layer, a few more ReLU gates flip downstream and the golden-size
  // conv1_1.weight gradient moved from 6e-3 to 1.3e-2 rel-L2, past SURVEY 8d's 1e-2.  SZN_CONV1_1_SIMT=1 forces it everywhere.
  static int simt = -1;
  if (simt < 0) simt = getenv("SZN_CONV1_1_SIMT") ? 1 : 0;
  if (!simt && dtype != SZN_F32X3) return conv1_1_fwd_tc(dtype, x, w_oihw, bias, y, B, H, W, pad, (cudaStream_t)stream);
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  const long long total = (long long)B * Ho * Wo;
  const unsigned grid = (unsigned)((total + 127) / 128);
  DISPATCH_T(dtype, (conv1_1_fwd_kernel<T><<<grid, 128, 0, (cudaStream_t)stream>>>(x, w_oihw, bias, y, B, H, W, Ho, Wo, pad)));
  return check_launch("szn_conv1_1_fwd");
}

extern "C" int szn_conv1_1_wgrad(int dtype, const float* x, const void* dy, float* dw_oihw, int B, int H, int W, int pad,
                                 void* stream) {
  const int Ho = H + 2 * pad - 2, Wo = W + 2 * pad - 2;
  const int ylo = pad - 2 > 0 ? pad - 2 : 0, xlo = ylo;
  const int yhi = pad + H - 1 < Ho - 1 ? pad + H - 1 : Ho - 1, xhi = pad + W - 1 < Wo - 1 ? pad + W - 1 : Wo - 1;
  const int wy = yhi - ylo + 1, wx = xhi - xlo + 1;  // output pixels whose window touches the image
  const int nseg = (wx + C11_SEG - 1) / C11_SEG;
  const long long items = (long long)B * wy * nseg;
  if (!env_flag("SZN_CONV1_1_WGRAD_V2", SZN_NEW_KERNELS_DEFAULT)) {  // the (co, ci, filter row) blocking (switch read per call: A/B runs, tests)
    const int grid = (int)(items < 148 * 3 ? items : 148 * 3);
    DISPATCH_T(dtype, (conv1_1_wgrad_kernel<T><<<grid, 576, 0, (cudaStream_t)stream>>>(x, dy, dw_oihw, B, H, W, Ho, Wo, pad, ylo, xlo, wy, nseg)));
    return check_launch("szn_conv1_1_wgrad");
  }
  static int per_sm[3] = {0, 0, 0}, sms = 0;  // resident CTAs per SM of each instantiation (persistent grid)
  if (dtype < 0 || dtype > 2) return set_error(SZN_ERR_ARG, "bad dtype");
  if (!per_sm[dtype]) {
    int dev = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    DISPATCH_T(dtype, (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, conv1_1_wgrad_v2_kernel<T>, C11_NT, 0)));
    per_sm[dtype] = occ > 0 ? occ : 1;
  }
  const long long cap = (long long)sms * per_sm[dtype];
  const int grid = (int)(items < cap ? items : cap);
  DISPATCH_T(dtype, (conv1_1_wgrad_v2_kernel<T><<<grid, C11_NT, 0, (cudaStream_t)stream>>>(x, dy, dw_oihw, B, H, W, Ho, Wo, pad, ylo, xlo, wy, nseg)));
  return check_launch("szn_conv1_1_wgrad");
}

extern "C" int szn_pool_fwd(int dtype, const void* in, void* out, int B, int H, int W, int C, void* stream) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int vn = dtype == SZN_F32 ? 4 : 8;
  if (C % vn) return set_error(SZN_ERR_ARG, "szn_pool_fwd: C must be a multiple of 16 bytes");
  const long long total = (long long)B * Ho * Wo * (C / vn);
  DISPATCH_T(dtype, (pool_fwd_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, H, W, C, Ho, Wo)));
  return check_launch("szn_pool_fwd");
}

extern "C" int szn_pool_bwd(int dtype, const void* y, const void* dp, void* dy, int B, int H, int W, int C, int relu_gate,
                            float* dy_col_sum, void* stream) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int vn = dtype == SZN_F32 ? 4 : 8;
  if (C % vn) return set_error(SZN_ERR_ARG, "szn_pool_bwd: C must be a multiple of 16 bytes");
  const long long total = (long long)B * Ho * Wo * (C / vn);
  if (dy_col_sum && 256 % (C / vn)) return set_error(SZN_ERR_UNSUPPORTED, "szn_pool_bwd: fused channel sums need C/vector to divide 256");
  int grid = grid_for(total, 256);
  {
    static int cap = 0;  // with fused channel sums every block ends with C atomics: fewer, longer-lived blocks
    if (!cap) {
      const char* e = getenv("SZN_POOL_BLOCKS");
      cap = e ? atoi(e) : 148 * 4;
    }
    if (dy_col_sum && grid > cap) grid = cap;
  }
  DISPATCH_T(dtype, (pool_bwd_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(y, dp, dy, B, H, W, C, Ho, Wo, relu_gate, dy_col_sum)));
  return check_launch("szn_pool_bwd");
}

extern "C" int szn_pool_fwd_code(int dtype, const void* in, void* out, unsigned char* code, int B, int H, int W, int C,
                                 void* stream) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int vn = dtype == SZN_F32 ? 4 : 8;
  if (C % vn) return set_error(SZN_ERR_ARG, "szn_pool_fwd_code: C must be a multiple of 16 bytes");
  if (!code || (reinterpret_cast<uintptr_t>(code) & 7)) return set_error(SZN_ERR_ARG, "szn_pool_fwd_code: code must be an 8-byte aligned buffer");
  const long long total = (long long)B * Ho * Wo * (C / vn);
  DISPATCH_T(dtype, (pool_fwd_code_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(in, out, code, B, H, W, C, Ho, Wo)));
  return check_launch("szn_pool_fwd_code");
}

extern "C" int szn_pool_bwd_code(int dtype, const unsigned char* code, const void* dp, void* dy, int B, int H, int W, int C,
                                 int relu_gate, float* dy_col_sum, void* stream) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int vn = dtype == SZN_F32 ? 4 : 8;
  if (C % vn) return set_error(SZN_ERR_ARG, "szn_pool_bwd_code: C must be a multiple of 16 bytes");
  if (!code || (reinterpret_cast<uintptr_t>(code) & 7)) return set_error(SZN_ERR_ARG, "szn_pool_bwd_code: code must be an 8-byte aligned buffer");
  const long long total = (long long)B * Ho * Wo * (C / vn);
  if (dy_col_sum && 256 % (C / vn)) return set_error(SZN_ERR_UNSUPPORTED, "szn_pool_bwd_code: fused channel sums need C/vector to divide 256");
  int grid = grid_for(total, 256);
  if (dy_col_sum && grid > 148 * 4) grid = 148 * 4;  // every block ends with C atomics: fewer, longer-lived blocks (as szn_pool_bwd)
  DISPATCH_T(dtype, (pool_bwd_code_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(code, dp, dy, B, H, W, C, Ho, Wo, relu_gate, dy_col_sum)));
  return check_launch("szn_pool_bwd_code");
}

extern "C" int szn_bias_grad(int dtype, const void* dy, float* db, long long rows, int C, long long ld, void* stream) {
  const int vn = dtype == SZN_F32 ? 4 : 8;
  if (C % vn || ld % vn || (reinterpret_cast<uintptr_t>(dy) & 15))
    return set_error(SZN_ERR_ARG, "szn_bias_grad: C, ld and dy must be 16-byte aligned");
  long long rpb = (rows + 148 * 8 - 1) / (148 * 8);
  if (rpb < 64) rpb = 64;
  const unsigned grid = (unsigned)((rows + rpb - 1) / rpb);
  DISPATCH_T(dtype, (bias_grad_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(dy, db, rows, C, ld, rpb)));
  return check_launch("szn_bias_grad");
}

extern "C" int szn_pack_weight(int dtype, const float* w_oihw, void* out, int O, int I, int R, int S, int O_pad,
                               void* stream) {
  const long long total = (long long)O_pad * R * S * I;
  DISPATCH_T(dtype, (pack_weight_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, out, O, I, R * S, O_pad)));
  return check_launch("szn_pack_weight");
}

extern "C" int szn_pack_weight_dgrad(int dtype, const float* w_oihw, void* out, int O, int I, int R, int S, int O_pad,
                                     int mode, void* stream) {
  const long long total = (long long)I * R * S * O_pad;
  DISPATCH_T(dtype, (pack_weight_dgrad_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, out, O, I, R, S, O_pad, mode)));
  return check_launch("szn_pack_weight_dgrad");
}

extern "C" int szn_col2im(int dtype, const void* dcol, void* dx, int B, int H, int W, int C, int R, int S, void* stream) {
  const int vn = dtype == SZN_F32 ? 4 : 8;
  if (C % vn) return set_error(SZN_ERR_ARG, "szn_col2im: C must be a multiple of 16 bytes");
  const long long total = (long long)B * H * W * (C / vn);
  DISPATCH_T(dtype, (col2im_kernel<T><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dcol, dx, B, H, W, C, R, S)));
  return check_launch("szn_col2im");
}

extern "C" int szn_unpack_wgrad(const float* dw_ohwi, float* g_oihw, int O, int I, int R, int S, void* stream) {
  const long long total = (long long)O * R * S * I;
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dw_ohwi, g_oihw, O, I, R * S);
  return check_launch("szn_unpack_wgrad");
}

extern "C" int szn_dropout_scale(float* scale, int n, unsigned long long seed, void* stream) {
  dropout_scale_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scale, n, seed);
  return check_launch("szn_dropout_scale");
}

extern "C" int szn_sgd_step(float* param, const float* grad, float* momentum_buf, long long n, float lr, float momentum,
                            float weight_decay, int first_step, void* stream) {
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(momentum_buf)) & 15)
    return set_error(SZN_ERR_ARG, "szn_sgd_step: buffers must be 16-byte aligned");
  sgd_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, momentum_buf, n, lr, momentum,
                                                                        weight_decay, first_step);
  return check_launch("szn_sgd_step");
}

extern "C" int szn_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float beta1,
                             float beta2, float eps, float weight_decay, float step_size, float bias_correction2_sqrt,
                             void* stream) {
  if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
       reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
    return set_error(SZN_ERR_ARG, "szn_adam_step: buffers must be 16-byte aligned");
  if (!(bias_correction2_sqrt > 0.f)) return set_error(SZN_ERR_ARG, "szn_adam_step: bias_correction2_sqrt must be > 0");
  adam_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, beta1, beta2,
                                                                         eps, weight_decay, step_size,
                                                                         bias_correction2_sqrt);
  return check_launch("szn_adam_step");
}

extern "C" int szn_cast(int dtype, const float* in, void* out, long long rows, int C, void* stream) {
  const long long n = rows * C;
  DISPATCH_T(dtype, (cast_kernel<T><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(in, out, n, C)));
  return check_launch("szn_cast");
}
