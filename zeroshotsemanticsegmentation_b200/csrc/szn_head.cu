// Head of the SZN path: x32 bilinear "upscore" + crop (models.py:94,146-151), the per-pixel losses
// (utils.py:19-102) and nearest-class-embedding inference (utils.py:159-205).
// These stages are HBM-bound: one coalesced pass over the (B,D,H,W) score tensor each.
#include "szn_internal.h"
#include "szn_ptx.cuh"
#include "szn_store.cuh"

namespace szn {

constexpr int UPK = 64, UPS = 32, UPCROP = 19;

template <int VEC>
struct PixVec {
  float v[VEC];
};
template <int VEC>
__device__ __forceinline__ PixVec<VEC> ld_pix(const float* p) {
  PixVec<VEC> r;
  if (VEC == 4) {
    const float4 f = __ldcs(reinterpret_cast<const float4*>(p));
    r.v[0] = f.x, r.v[1 % VEC] = f.y, r.v[2 % VEC] = f.z, r.v[3 % VEC] = f.w;
  } else {
    r.v[0] = __ldcs(p);
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void st_pix(float* p, const PixVec<VEC>& r) {
  if (VEC == 4) __stcs(reinterpret_cast<float4*>(p), make_float4(r.v[0], r.v[1 % VEC], r.v[2 % VEC], r.v[3 % VEC]));
  else __stcs(p, r.v[0]);
}



// 1-D tent of get_upsampling_weight (models.py:11-19): k = 64 -> factor 32, centre 31.5
__device__ __forceinline__ float tent(int k) { return 1.f - fabsf((float)k - 31.5f) * (1.f / 32.f); }

// ------------------------------------------------------------------------------------------------
// upscore with the diagonal bilinear weight == per-channel x32 bilinear upsample, cropped at 19.
// s: [B,hs,ws,ld] fp32 (channels coff..coff+D) -> out: NCHW fp32 [B,D,H,W].
// HBM-bound on the write of out (B*D*H*W*4 bytes).  One CTA = one (b, d) plane x UP_ROWS output rows: the plane's
// hs x ws source values sit zero-padded in shared memory (no border branches), each thread produces VEC consecutive
// X per row (16-byte coalesced stores) from 3 vertically interpolated source columns.
// ------------------------------------------------------------------------------------------------
constexpr int UP_ROWS = 64;
constexpr int UP_MAXSRC = 34;  // hs, ws <= 32 (inputs up to ~900 px); larger maps use more shared memory than we reserve

template <int VEC>
__global__ void __launch_bounds__(128) upsample_fwd_kernel(const float* __restrict__ s, float* __restrict__ out, int B, int D,
                                                           int H, int W, int hs, int ws, int ld, int coff) {
  __shared__ float sp[UP_MAXSRC][UP_MAXSRC + 1];  // sp[i+1][j+1] = s[i][j], zero border
  const int plane = blockIdx.x;                   // b * D + d
  const int b = plane / D, d = plane - b * D;
  const int y_begin = blockIdx.y * UP_ROWS;
  for (int i = threadIdx.x; i < (hs + 2) * (ws + 2); i += blockDim.x) {
    const int r = i / (ws + 2), c = i - r * (ws + 2);
    float v = 0.f;
    if (r >= 1 && r <= hs && c >= 1 && c <= ws) v = s[(((long long)b * hs + (r - 1)) * ws + (c - 1)) * ld + coff + d];
    sp[r][c] = v;
  }
  __syncthreads();
  float* oplane = out + (long long)plane * H * W;
  const int wv = (W + VEC - 1) / VEC;
  for (int xq = threadIdx.x; xq < wv; xq += blockDim.x) {
    // horizontal taps of this thread's VEC pixels: source columns ix1 (weight fx1) and ix1-1 (weight fx0)
    int cb = ((xq * VEC + UPCROP) >> 5);  // padded index of column ix1-1 for the first pixel
    float fx1[VEC], fx0[VEC];
    int sel[VEC];  // 0: columns (cb, cb+1), 1: columns (cb+1, cb+2)
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const int v = xq * VEC + j + UPCROP;
      sel[j] = (v >> 5) - cb;
      fx1[j] = tent(v & 31), fx0[j] = tent((v & 31) + 32);
    }
    int y_end = y_begin + UP_ROWS;
    if (y_end > H) y_end = H;
    for (int Y = y_begin; Y < y_end; ++Y) {
      const int u = Y + UPCROP;
      const int r0 = u >> 5;  // padded row index of iy1-1; iy1 is r0+1
      const float fy1 = tent(u & 31), fy0 = tent((u & 31) + 32);
      const float v0 = fy1 * sp[r0 + 1][cb] + fy0 * sp[r0][cb];
      const float v1 = fy1 * sp[r0 + 1][cb + 1] + fy0 * sp[r0][cb + 1];
      const float v2 = fy1 * sp[r0 + 1][cb + 2] + fy0 * sp[r0][cb + 2];
      float o[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = sel[j] ? (fx1[j] * v2 + fx0[j] * v1) : (fx1[j] * v1 + fx0[j] * v0);
      float* dst = oplane + (long long)Y * W + xq * VEC;
      if (VEC == 4) {
        __stcs(reinterpret_cast<float4*>(dst), make_float4(o[0], o[1], o[2], o[3]));
      } else {
        dst[0] = o[0];
      }
    }
  }
}

// transpose of the above: ds[b,iy,ix,coff+d] = sum_{ky,kx} g[b,d,32iy+ky-19,32ix+kx-19] f(ky) f(kx).
// HBM-bound on the read of g.  One CTA per (b, d) plane streams its H rows ONCE (coalesced, 8 rows in flight per
// thread): a row feeds source row iy1 with weight f(ky) and iy1-1 with f(ky+32); when a 32-row group ends, the finished
// column sums are folded along x by 8 threads per ix (64 taps each).
// VEC = 4: a thread owns 4 adjacent columns and reads 16 bytes per row (W % 4 == 0); VEC = 1: any W.
// NS column slots per thread (slot j = columns (threadIdx.x + j * blockDim.x) * VEC ...), so W <= blockDim.x * VEC * NS:
// native-size PASCAL images (e.g. 500 x 375, train.py:82-84 feeds them unpadded at batch size 1) take VEC = 1, NS = 2.
template <typename T, int VEC, int NS>
__global__ void __launch_bounds__(256) upsample_bwd_kernel(const float* __restrict__ g, void* __restrict__ ds, int B, int D,
                                                           int H, int W, int hs, int ws, int ld, int coff) {
  extern __shared__ float col[];  // [W]
  const int plane = blockIdx.x;
  const int b = plane / D, d = plane - b * D;
  const float* gp = g + (long long)plane * H * W;
  float carry[NS][VEC], accA[NS][VEC], accB[NS][VEC];
#pragma unroll
  for (int j = 0; j < NS; ++j)
#pragma unroll
    for (int c = 0; c < VEC; ++c) carry[j][c] = accA[j][c] = accB[j][c] = 0.f;
  const int n_groups = (H + UPCROP + 31) >> 5;  // groups of rows with the same iy1 = (Y + 19) >> 5
  for (int k = 0; k <= n_groups; ++k) {
    if (k < n_groups) {  // rows of group k: u = Y + 19 in [32k, 32k + 31]
      int y0 = 32 * k - UPCROP, y1 = y0 + 32;
      if (y0 < 0) y0 = 0;
      if (y1 > H) y1 = H;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int X0 = (threadIdx.x + j * blockDim.x) * VEC;
        if (X0 >= W) continue;
#pragma unroll 8
        for (int Y = y0; Y < y1; ++Y) {
          const int ky = (Y + UPCROP) & 31;
          const float fy1 = tent(ky), fy0 = tent(ky + 32);
          const PixVec<VEC> v = ld_pix<VEC>(gp + (long long)Y * W + X0);
#pragma unroll
          for (int c = 0; c < VEC; ++c) {
            accA[j][c] = fmaf(v.v[c], fy1, accA[j][c]);
            accB[j][c] = fmaf(v.v[c], fy0, accB[j][c]);
          }
        }
      }
    }
    // source row iy = k - 1 is complete: carry (its f(ky) part from group k-1) + accB (its f(ky+32) part from group k)
    const int iy = k - 1;
    if (iy >= 0 && iy < hs) {
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const int X0 = (threadIdx.x + j * blockDim.x) * VEC;
        if (X0 < W) {
#pragma unroll
          for (int c = 0; c < VEC; ++c) col[X0 + c] = carry[j][c] + accB[j][c];
        }
      }
      __syncthreads();
      const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7, ngrp = blockDim.x >> 3;  // groups of 8 threads
      for (int ix0 = 0; ix0 < ws; ix0 += ngrp) {  // warp-uniform trip count: every lane takes part in the shuffles
        const int ix = ix0 + grp;
        float acc = 0.f;
        if (ix < ws) {
          for (int kx = sub; kx < UPK; kx += 8) {
            const int X = UPS * ix + kx - UPCROP;
            if (X >= 0 && X < W) acc = fmaf(col[X], tent(kx), acc);
          }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (sub == 0 && ix < ws) Store<T>::store_elem(ds, ((long long)b * hs + iy) * ws + ix, ld, coff + d, acc);
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < NS; ++j)
#pragma unroll
      for (int c = 0; c < VEC; ++c) carry[j][c] = accA[j][c], accA[j][c] = 0.f, accB[j][c] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// dense ConvTranspose2d(Ci, Co, 64, stride 32) + crop for SMALL channel counts (the trained 2x2 seenmask
// head, models.py:98,150-151; also the fallback when upscore.weight is not the diagonal bilinear filter).
// wd: [Ci][Co][64][64] fp32 (PyTorch layout).
// ------------------------------------------------------------------------------------------------
__global__ void deconv_fwd_kernel(const float* __restrict__ s, const float* __restrict__ wd, float* __restrict__ out,
                                  int B, int Ci, int Co, int H, int W, int hs, int ws, int ld, int coff) {
  const long long total = (long long)B * Co * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(i % W);
    long long r = i / W;
    const int Y = (int)(r % H);
    r /= H;
    const int j = (int)(r % Co);
    const int b = (int)(r / Co);
    const int u = Y + UPCROP, v = X + UPCROP;
    const int iy1 = u >> 5, ky1 = u & 31, ix1 = v >> 5, kx1 = v & 31;
    float acc = 0.f;
    for (int ci = 0; ci < Ci; ++ci) {
      const float* wp = wd + ((long long)ci * Co + j) * UPK * UPK;
      const float* sp = s + (long long)b * hs * ws * ld + coff + ci;
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const int iy = iy1 - a, ky = ky1 + 32 * a;
        if (iy < 0 || iy >= hs) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int ix = ix1 - c, kx = kx1 + 32 * c;
          if (ix < 0 || ix >= ws) continue;
          acc = fmaf(sp[((long long)iy * ws + ix) * ld], __ldg(wp + ky * UPK + kx), acc);
        }
      }
    }
    out[i] = acc;
  }
}

// ds[b,iy,ix,coff+ci] = sum_{j,ky,kx} g[b,j,32iy+ky-19,32ix+kx-19] wd[ci][j][ky][kx]   (one CTA per (b,iy,ix))
template <typename T>
__global__ void __launch_bounds__(256) deconv_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ wd,
                                                           void* __restrict__ ds, int B, int Ci, int Co, int H, int W,
                                                           int hs, int ws, int ld, int coff) {
  __shared__ float red[8];
  const int ix = blockIdx.x % ws, iy = (blockIdx.x / ws) % hs, b = blockIdx.x / (ws * hs);
  for (int ci = 0; ci < Ci; ++ci) {
    float acc = 0.f;
    for (int j = 0; j < Co; ++j) {
      const float* gp = g + ((long long)b * Co + j) * H * W;
      const float* wp = wd + ((long long)ci * Co + j) * UPK * UPK;
      for (int t = threadIdx.x; t < UPK * UPK; t += 256) {
        const int ky = t >> 6, kx = t & 63;
        const int Y = UPS * iy + ky - UPCROP, X = UPS * ix + kx - UPCROP;
        if (Y >= 0 && Y < H && X >= 0 && X < W) acc = fmaf(gp[(long long)Y * W + X], __ldg(wp + t), acc);
      }
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int k = 0; k < 8; ++k) t += red[k];
      Store<T>::store_elem(ds, ((long long)b * hs + iy) * ws + ix, ld, coff + ci, t);
    }
    __syncthreads();
  }
}

// dwd[ci][j][ky][kx] = sum_{b,iy,ix} s[b,iy,ix,coff+ci] g[b,j,32iy+ky-19,32ix+kx-19]
__global__ void deconv_wgrad_kernel(const float* __restrict__ s, const float* __restrict__ g, float* __restrict__ dwd,
                                    int B, int Ci, int Co, int H, int W, int hs, int ws, int ld, int coff) {
  const long long total = (long long)Ci * Co * UPK * UPK;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kx = (int)(i & 63), ky = (int)((i >> 6) & 63);
    const int j = (int)((i >> 12) % Co), ci = (int)((i >> 12) / Co);
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      const float* gp = g + ((long long)b * Co + j) * H * W;
      const float* sp = s + (long long)b * hs * ws * ld + coff + ci;
      for (int iy = 0; iy < hs; ++iy) {
        const int Y = UPS * iy + ky - UPCROP;
        if (Y < 0 || Y >= H) continue;
        for (int ix = 0; ix < ws; ++ix) {
          const int X = UPS * ix + kx - UPCROP;
          if (X < 0 || X >= W) continue;
          acc = fmaf(sp[((long long)iy * ws + ix) * ld], gp[(long long)Y * W + X], acc);
        }
      }
    }
    dwd[i] = acc;
  }
}

// ------------------------------------------------------------------------------------------------
// per-pixel losses on the NCHW fp32 score tensor.  One thread per pixel, loop over channels: each
// channel plane is read with fully coalesced warp accesses.  accum = {sum, n_valid} in fp64.
// kind: 0 cosine (utils.py:75-102), 1 mse (utils.py:50-73)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_accumulate(double a, double b, double* accum) {
  __shared__ double ra[8], rb[8];
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) ra[w] = a, rb[w] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0, sb = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) sa += ra[k], sb += rb[k];
    atomicAdd(accum, sa);
    atomicAdd(accum + 1, sb);
  }
}

// One thread = VEC consecutive pixels of one image (VEC = 4: 16-byte loads of every channel plane, fully coalesced),
// channel loop unrolled 4x so 4-8 independent 16-byte loads per thread are in flight: HBM-bound, one pass over score.
template <int KIND, int VEC, bool HAS_TE>
__global__ void __launch_bounds__(256, 3) embed_loss_fwd_kernel(const float* __restrict__ score, const long long* __restrict__ target,
                                                             const float* __restrict__ te, const float* __restrict__ table,
                                                             int n, int c, long long hw, int rows, float* __restrict__ stats,
                                                             double* __restrict__ accum) {
  // rows = rows of `table`: a label >= rows (an un-remapped 255, a table of another dataset) must not index past it.
  // torch's embedding / nll_loss device-assert on such input; here the pixel reads row rows-1 and poisons the loss with
  // NaN, which the trainers' NaN guard (trainer_fcn.py:107-108) turns into an exception.
  const long long total = n * hw / VEC;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double part = 0, cnt = 0;
  if (i < total) {
    const long long pix = i * VEC;  // hw % VEC == 0: the VEC pixels share one image
    const long long b = pix / hw, p = pix - b * hw;
    long long t[VEC];
    bool any = false;
#pragma unroll
    for (int j = 0; j < VEC; ++j) t[j] = target[pix + j], any |= t[j] >= 0;
    if (any) {
      const float* sp = score + b * c * hw + p;
      const float* ep = HAS_TE ? te + b * c * hw + p : nullptr;
      const float* tp[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) tp[j] = HAS_TE ? nullptr : table + (t[j] < 0 ? 0 : t[j] < rows ? t[j] : rows - 1) * c;
      float ss[VEC], se[VEC], ee[VEC], sq[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) ss[j] = se[j] = ee[j] = sq[j] = 0.f;
#pragma unroll 8
      for (int d = 0; d < c; ++d) {
        const PixVec<VEC> sv = ld_pix<VEC>(sp + d * hw);
        PixVec<VEC> ev;
        if (HAS_TE) {
          ev = ld_pix<VEC>(ep + d * hw);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) ev.v[j] = __ldg(tp[j] + d);
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          if (KIND == 0) {
            ss[j] = fmaf(sv.v[j], sv.v[j], ss[j]);
            se[j] = fmaf(sv.v[j], ev.v[j], se[j]);
            ee[j] = fmaf(ev.v[j], ev.v[j], ee[j]);
          } else {
            const float df = sv.v[j] - ev.v[j];
            sq[j] = fmaf(df, df, sq[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        if (t[j] < 0) continue;
        if (!HAS_TE && t[j] >= rows) part += (double)NAN;  // out-of-range label
        if (KIND == 0) {
          const float inv_s = 1.f / sqrtf(ss[j]), inv_e = 1.f / sqrtf(ee[j]);
          const float cs = se[j] * inv_s * inv_e;
          float* st = stats + 3 * (pix + j);
          st[0] = inv_s, st[1] = inv_e, st[2] = cs;
          part += cs;
        } else {
          part += sq[j];
        }
        cnt += 1;
      }
    }
  }
  block_accumulate(part, cnt, accum);
}

template <int KIND, int VEC, bool HAS_TE>
__global__ void __launch_bounds__(256, 2) embed_loss_bwd_kernel(const float* __restrict__ score, const long long* __restrict__ target,
                                                             const float* __restrict__ te, const float* __restrict__ table,
                                                             int n, int c, long long hw, int rows, const float* __restrict__ stats,
                                                             const double* __restrict__ accum, const float* __restrict__ gout,
                                                             float* __restrict__ dscore) {
  const long long total = n * hw / VEC;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long pix = i * VEC;
  const long long b = pix / hw, p = pix - b * hw;
  long long t[VEC];
  bool any = false;
#pragma unroll
  for (int j = 0; j < VEC; ++j) t[j] = target[pix + j], any |= t[j] >= 0;
  float* gp = dscore + b * c * hw + p;
  if (!any) {
    PixVec<VEC> z;
#pragma unroll
    for (int j = 0; j < VEC; ++j) z.v[j] = 0.f;
    for (int d = 0; d < c; ++d) st_pix<VEC>(gp + d * hw, z);
    return;
  }
  const float gs = gout[0] / (float)accum[1];
  const float* sp = score + b * c * hw + p;
  const float* ep = HAS_TE ? te + b * c * hw + p : nullptr;
  const float* tp[VEC];
  float inv_s[VEC], inv_e[VEC], cs[VEC], live[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    tp[j] = HAS_TE ? nullptr : table + (t[j] < 0 ? 0 : t[j] < rows ? t[j] : rows - 1) * c;
    live[j] = t[j] >= 0 ? 1.f : 0.f;
    inv_s[j] = inv_e[j] = cs[j] = 0.f;
    if (KIND == 0 && t[j] >= 0) {
      const float* st = stats + 3 * (pix + j);
      inv_s[j] = st[0], inv_e[j] = st[1], cs[j] = st[2];
    }
  }
  // Explicit batches of 8 channels: all 8 (or 16) vector loads are issued before the first store.  Written as one
  // load-compute-store per channel, the cache-hinted loads were kept in program order behind the previous channel's
  // store (one 16-byte load in flight per thread: 4.1 TB/s).
  constexpr int UNR = 8;
  for (int d0 = 0; d0 < c; d0 += UNR) {
    PixVec<VEC> sv[UNR], ev[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int d = d0 + u < c ? d0 + u : c - 1;  // the tail re-reads the last channel (never stored twice)
      sv[u] = ld_pix<VEC>(sp + d * hw);
      if (HAS_TE) {
        ev[u] = ld_pix<VEC>(ep + d * hw);
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) ev[u].v[j] = __ldg(tp[j] + d);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if (d0 + u >= c) break;
      PixVec<VEC> gv;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float g;
        if (KIND == 0) g = -gs * (ev[u].v[j] * inv_e[j] - cs[j] * sv[u].v[j] * inv_s[j]) * inv_s[j];  // d(-cos)/ds
        else g = 2.f * gs * (sv[u].v[j] - ev[u].v[j]);
        gv.v[j] = live[j] != 0.f ? g : 0.f;
      }
      st_pix<VEC>(gp + (d0 + u) * hw, gv);
    }
  }
}

// cross_entropy2d (utils.py:19-48): log-softmax over c, NLL summed over target >= 0
__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ score, const long long* __restrict__ target,
                                                     int n, int c, long long hw, float* __restrict__ lse_out,
                                                     double* __restrict__ accum) {
  const long long total = n * hw;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double part = 0, cnt = 0;
  if (i < total) {
    const long long t = target[i];
    const long long b = i / hw, p = i - b * hw;
    const float* sp = score + b * c * hw + p;
    float m = -INFINITY;
    for (int d = 0; d < c; ++d) m = fmaxf(m, sp[d * hw]);
    float se = 0.f;
    for (int d = 0; d < c; ++d) se += expf(sp[d * hw] - m);
    const float lse = m + logf(se);
    lse_out[i] = lse;
    if (t >= c) {
      part = (double)NAN, cnt = 1;  // out-of-range label: poison the loss instead of reading past the score (see above)
    } else if (t >= 0) {
      part = (double)(lse - sp[t * hw]);
      cnt = 1;
    }
  }
  block_accumulate(part, cnt, accum);
}

__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ score, const long long* __restrict__ target,
                                                     int n, int c, long long hw, const float* __restrict__ lse,
                                                     const double* __restrict__ accum, const float* __restrict__ gout,
                                                     int size_average, float* __restrict__ dscore) {
  const long long total = n * hw;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long t = target[i];
  const long long b = i / hw, p = i - b * hw;
  const float* sp = score + b * c * hw + p;
  float* gp = dscore + b * c * hw + p;
  if (t < 0 || t >= c) {
    for (int d = 0; d < c; ++d) gp[d * hw] = 0.f;
    return;
  }
  const float gs = size_average ? gout[0] / (float)accum[1] : gout[0];
  const float l = lse[i];
  for (int d = 0; d < c; ++d) {
    const float pr = expf(sp[d * hw] - l);
    gp[d * hw] = gs * (pr - (d == t ? 1.f : 0.f));
  }
}

// kind 0: (N - sum)/N   1: sum/N   2: sum   3: sum/N  (N = accum[1])
__global__ void loss_finalize_kernel(const double* accum, int kind, float* loss) {
  const double s = accum[0], n = accum[1];
  double v;
  if (kind == 0) v = (n - s) / n;
  else if (kind == 2) v = s;
  else v = s / n;
  loss[0] = (float)v;
}

// ------------------------------------------------------------------------------------------------
// infer_lbl (utils.py:159-185): labels[p] = argmax_c  <s_p, e_c> / (|s_p| * |e_c|) with |e_c| == 0 -> 1,
// first index on ties.  fp32 CUDA-core contraction: CTA = 128 pixels x 64-class chunks, 4x8 register tile.
// ------------------------------------------------------------------------------------------------
__global__ void table_norm_kernel(const float* __restrict__ table, int C, int D, float* __restrict__ en) {
  const int cidx = blockIdx.x * blockDim.x + threadIdx.x;
  if (cidx >= C) return;
  float ss = 0.f;
  for (int d = 0; d < D; ++d) ss = fmaf(table[(long long)cidx * D + d], table[(long long)cidx * D + d], ss);
  const float nrm = sqrtf(ss);
  en[cidx] = nrm == 0.f ? 1.f : nrm;
}

__global__ void __launch_bounds__(256) embed_argmax_kernel(const float* __restrict__ score, const float* __restrict__ table,
                                                           const float* __restrict__ en, int n, int D, long long hw, int C,
                                                           long long* __restrict__ labels) {
  constexpr int KD = 16;
  __shared__ float ss_[KD][128];
  __shared__ float st_[KD][64 + 1];
  __shared__ float bval[8][128];
  __shared__ int bidx[8][128];
  const int pg = threadIdx.x & 31, cg = threadIdx.x >> 5;  // 4 pixels x 8 classes per thread
  const long long tiles_per_img = (hw + 127) / 128;
  const long long b = blockIdx.x / tiles_per_img;
  const long long p0 = (blockIdx.x % tiles_per_img) * 128;
  const float* sp = score + b * D * hw;
  float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int besti[4] = {0, 0, 0, 0};
  float nrm2[4] = {0, 0, 0, 0};
  for (int c0 = 0; c0 < C; c0 += 64) {
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[a][k] = 0.f;
    for (int d0 = 0; d0 < D; d0 += KD) {
      for (int i = threadIdx.x; i < KD * 128; i += 256) {
        const int dd = i >> 7, px = i & 127;
        const long long p = p0 + px;
        ss_[dd][px] = (d0 + dd < D && p < hw) ? sp[(long long)(d0 + dd) * hw + p] : 0.f;
      }
      for (int i = threadIdx.x; i < KD * 64; i += 256) {
        const int cc = i / KD, dd = i - cc * KD;
        st_[dd][cc] = (c0 + cc < C && d0 + dd < D) ? __ldg(table + (long long)(c0 + cc) * D + d0 + dd) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int dd = 0; dd < KD; ++dd) {
        float sv[4], tv[8];
#pragma unroll
        for (int a = 0; a < 4; ++a) sv[a] = ss_[dd][pg * 4 + a];
#pragma unroll
        for (int k = 0; k < 8; ++k) tv[k] = st_[dd][cg * 8 + k];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          if (c0 == 0) nrm2[a] = fmaf(sv[a], sv[a], nrm2[a]);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[a][k] = fmaf(sv[a], tv[k], acc[a][k]);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float sn = sqrtf(nrm2[a]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int cls = c0 + cg * 8 + k;
        if (cls < C) {
          const float v = acc[a][k] / (sn * __ldg(en + cls));
          if (v > best[a]) best[a] = v, besti[a] = cls;
        }
      }
    }
  }
  // nrm2 was only accumulated during the first class chunk; later chunks reuse it (same pixels)
#pragma unroll
  for (int a = 0; a < 4; ++a) bval[cg][pg * 4 + a] = best[a], bidx[cg][pg * 4 + a] = besti[a];
  __syncthreads();
  if (threadIdx.x < 128) {
    const long long p = p0 + threadIdx.x;
    if (p < hw) {
      float bv = bval[0][threadIdx.x];
      int bi = bidx[0][threadIdx.x];
      for (int k = 1; k < 8; ++k) {
        const float v = bval[k][threadIdx.x];
        const int ii = bidx[k][threadIdx.x];
        if (v > bv || (v == bv && ii < bi)) bv = v, bi = ii;
      }
      labels[b * hw + p] = bi;
    }
  }
}

// stich_seen_unseen_with_mask (utils.py:201-205): out = unseen_mask ? lbl_unseen : lbl_seen, where the mask is
//   mode 0: 1 - argmax(seen_mask_score[n,2,h,w])  (utils.py:195-199; ties -> channel 0 -> "unseen")
//   mode 1: target label is in the unseen list (utils.py:188-192)
__global__ void stitch_kernel(const long long* __restrict__ lbl_seen, const long long* __restrict__ lbl_unseen,
                              const float* __restrict__ sm, const long long* __restrict__ target,
                              const long long* __restrict__ unseen, int n_unseen, int mode, int n, long long hw,
                              long long* __restrict__ out) {
  const long long total = n * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    bool um;
    if (mode == 0) {
      const long long b = i / hw, p = i - b * hw;
      const float s0 = sm[(b * 2) * hw + p], s1 = sm[(b * 2 + 1) * hw + p];
      um = !(s1 > s0);
    } else {
      const long long t = target[i];
      um = false;
      for (int k = 0; k < n_unseen; ++k) um |= (t == unseen[k]);
    }
    out[i] = um ? lbl_unseen[i] : lbl_seen[i];
  }
}

static int hgrid(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace szn
using namespace szn;

extern "C" int szn_upsample32_crop_fwd(const float* s, float* out, int B, int D, int H, int W, int hs, int ws, int ld,
                                       int coff, void* stream) {
  if (hs + 2 > UP_MAXSRC || ws + 2 > UP_MAXSRC) return set_error(SZN_ERR_UNSUPPORTED, "szn_upsample32_crop_fwd: map too large");
  const dim3 grid((unsigned)((long long)B * D), (unsigned)((H + UP_ROWS - 1) / UP_ROWS));
  if (W % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
    upsample_fwd_kernel<4><<<grid, 128, 0, (cudaStream_t)stream>>>(s, out, B, D, H, W, hs, ws, ld, coff);
  else
    upsample_fwd_kernel<1><<<grid, 128, 0, (cudaStream_t)stream>>>(s, out, B, D, H, W, hs, ws, ld, coff);
  return check_launch("szn_upsample32_crop_fwd");
}

template <typename T>
static void launch_upsample_bwd(const float* g, void* ds, int B, int D, int H, int W, int hs, int ws, int ld, int coff,
                                bool v4, cudaStream_t st) {
  const unsigned grid = (unsigned)((long long)B * D);
  const size_t smem = (size_t)W * sizeof(float);
  if (v4) {
    int threads = (W / 4 + 31) / 32 * 32;
    if (threads < 64) threads = 64;
    upsample_bwd_kernel<T, 4, 1><<<grid, threads, smem, st>>>(g, ds, B, D, H, W, hs, ws, ld, coff);
  } else if (W <= 256) {
    upsample_bwd_kernel<T, 1, 1><<<grid, 256, smem, st>>>(g, ds, B, D, H, W, hs, ws, ld, coff);
  } else if (W <= 512) {
    upsample_bwd_kernel<T, 1, 2><<<grid, 256, smem, st>>>(g, ds, B, D, H, W, hs, ws, ld, coff);
  } else {
    upsample_bwd_kernel<T, 1, 4><<<grid, 256, smem, st>>>(g, ds, B, D, H, W, hs, ws, ld, coff);
  }
}

extern "C" int szn_upsample32_crop_bwd(int dtype, const float* g, void* ds, int B, int D, int H, int W, int hs, int ws,
                                       int ld, int coff, void* stream) {
  const bool v4 = W % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  if (W > 1024) return set_error(SZN_ERR_UNSUPPORTED, "szn_upsample32_crop_bwd: W > 1024");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SZN_BF16) launch_upsample_bwd<__nv_bfloat16>(g, ds, B, D, H, W, hs, ws, ld, coff, v4, st);
  else if (dtype == SZN_F32X3) launch_upsample_bwd<SplitBf16>(g, ds, B, D, H, W, hs, ws, ld, coff, v4, st);
  else launch_upsample_bwd<float>(g, ds, B, D, H, W, hs, ws, ld, coff, v4, st);
  return check_launch("szn_upsample32_crop_bwd");
}

extern "C" int szn_deconv_small_fwd(const float* s, const float* wd, float* out, int B, int Ci, int Co, int H, int W,
                                    int hs, int ws, int ld, int coff, void* stream) {
  const long long total = (long long)B * Co * H * W;
  deconv_fwd_kernel<<<hgrid(total, 256), 256, 0, (cudaStream_t)stream>>>(s, wd, out, B, Ci, Co, H, W, hs, ws, ld, coff);
  return check_launch("szn_deconv_small_fwd");
}

extern "C" int szn_deconv_small_dgrad(int dtype, const float* g, const float* wd, void* ds, int B, int Ci, int Co, int H,
                                      int W, int hs, int ws, int ld, int coff, void* stream) {
  const unsigned grid = (unsigned)(B * hs * ws);
  if (dtype == SZN_BF16)
    deconv_dgrad_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(g, wd, ds, B, Ci, Co, H, W, hs, ws, ld, coff);
  else if (dtype == SZN_F32X3)
    deconv_dgrad_kernel<SplitBf16><<<grid, 256, 0, (cudaStream_t)stream>>>(g, wd, ds, B, Ci, Co, H, W, hs, ws, ld, coff);
  else
    deconv_dgrad_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(g, wd, ds, B, Ci, Co, H, W, hs, ws, ld, coff);
  return check_launch("szn_deconv_small_dgrad");
}

extern "C" int szn_deconv_small_wgrad(const float* s, const float* g, float* dwd, int B, int Ci, int Co, int H, int W,
                                      int hs, int ws, int ld, int coff, void* stream) {
  const long long total = (long long)Ci * Co * UPK * UPK;
  deconv_wgrad_kernel<<<hgrid(total, 128), 128, 0, (cudaStream_t)stream>>>(s, g, dwd, B, Ci, Co, H, W, hs, ws, ld, coff);
  return check_launch("szn_deconv_small_wgrad");
}

// kind: 0 cosine, 1 mse.  Exactly one of target_embed ([n,c,h,w]) / table ([C,c], row = label, -1 ignored) is given.
// stats: [n*h*w*3] fp32 scratch kept for backward (cosine); accum: 2 doubles {sum, n_valid}; loss: 1 float.
extern "C" int szn_embed_loss_fwd(int kind, const float* score, const long long* target, const float* target_embed,
                                  const float* table, int table_rows, int n, int c, int h, int w, float* stats,
                                  double* accum, float* loss, void* stream) {
  if ((target_embed == nullptr) == (table == nullptr))
    return set_error(SZN_ERR_ARG, "szn_embed_loss_fwd: give exactly one of target_embed / table");
  if (table && table_rows < 1) return set_error(SZN_ERR_ARG, "szn_embed_loss_fwd: table_rows must be the number of table rows");
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)h * w, total = n * hw;
  cudaMemsetAsync(accum, 0, 2 * sizeof(double), st);
  if (kind != 0 && kind != 1) return set_error(SZN_ERR_ARG, "szn_embed_loss_fwd: kind");
  const bool v4 = hw % 4 == 0 && ((reinterpret_cast<uintptr_t>(score) | reinterpret_cast<uintptr_t>(target_embed)) & 15) == 0;
  const unsigned grid = (unsigned)((total / (v4 ? 4 : 1) + 255) / 256);
#define SZN_LOSS_FWD(K, V, TE) embed_loss_fwd_kernel<K, V, TE><<<grid, 256, 0, st>>>(score, target, target_embed, table, n, c, hw, table_rows, stats, accum)
  const bool has_te = target_embed != nullptr;
  if (kind == 0 && v4) { if (has_te) SZN_LOSS_FWD(0, 4, true); else SZN_LOSS_FWD(0, 4, false); }
  else if (kind == 0) { if (has_te) SZN_LOSS_FWD(0, 1, true); else SZN_LOSS_FWD(0, 1, false); }
  else if (v4) { if (has_te) SZN_LOSS_FWD(1, 4, true); else SZN_LOSS_FWD(1, 4, false); }
  else { if (has_te) SZN_LOSS_FWD(1, 1, true); else SZN_LOSS_FWD(1, 1, false); }
#undef SZN_LOSS_FWD
  if (int e = check_launch("szn_embed_loss_fwd")) return e;
  loss_finalize_kernel<<<1, 1, 0, st>>>(accum, kind, loss);
  return check_launch("szn_embed_loss_fwd/finalize");
}

extern "C" int szn_loss_finalize(int kind, const double* accum, float* loss, void* stream) {
  loss_finalize_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(accum, kind, loss);
  return check_launch("szn_loss_finalize");
}

extern "C" int szn_embed_loss_bwd(int kind, const float* score, const long long* target, const float* target_embed,
                                  const float* table, int table_rows, int n, int c, int h, int w, const float* stats,
                                  const double* accum, const float* grad_out, float* dscore, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)h * w, total = n * hw;
  if (kind != 0 && kind != 1) return set_error(SZN_ERR_ARG, "szn_embed_loss_bwd: kind");
  const bool v4 = hw % 4 == 0 && ((reinterpret_cast<uintptr_t>(score) | reinterpret_cast<uintptr_t>(target_embed) |
                                   reinterpret_cast<uintptr_t>(dscore)) & 15) == 0;
  const unsigned grid = (unsigned)((total / (v4 ? 4 : 1) + 255) / 256);
#define SZN_LOSS_BWD(K, V, TE) embed_loss_bwd_kernel<K, V, TE><<<grid, 256, 0, st>>>(score, target, target_embed, table, n, c, hw, table_rows, stats, accum, grad_out, dscore)
  const bool has_te = target_embed != nullptr;
  if (kind == 0 && v4) { if (has_te) SZN_LOSS_BWD(0, 4, true); else SZN_LOSS_BWD(0, 4, false); }
  else if (kind == 0) { if (has_te) SZN_LOSS_BWD(0, 1, true); else SZN_LOSS_BWD(0, 1, false); }
  else if (v4) { if (has_te) SZN_LOSS_BWD(1, 4, true); else SZN_LOSS_BWD(1, 4, false); }
  else { if (has_te) SZN_LOSS_BWD(1, 1, true); else SZN_LOSS_BWD(1, 1, false); }
#undef SZN_LOSS_BWD
  return check_launch("szn_embed_loss_bwd");
}

extern "C" int szn_ce2d_fwd(const float* score, const long long* target, int n, int c, int h, int w, int size_average,
                            float* lse, double* accum, float* loss, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const long long hw = (long long)h * w, total = n * hw;
  cudaMemsetAsync(accum, 0, 2 * sizeof(double), st);
  ce_fwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(score, target, n, c, hw, lse, accum);
  if (int e = check_launch("szn_ce2d_fwd")) return e;
  loss_finalize_kernel<<<1, 1, 0, st>>>(accum, size_average ? 3 : 2, loss);
  return check_launch("szn_ce2d_fwd/finalize");
}

extern "C" int szn_ce2d_bwd(const float* score, const long long* target, int n, int c, int h, int w, int size_average,
                            const float* lse, const double* accum, const float* grad_out, float* dscore, void* stream) {
  const long long hw = (long long)h * w, total = n * hw;
  ce_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(score, target, n, c, hw, lse, accum,
                                                                                     grad_out, size_average, dscore);
  return check_launch("szn_ce2d_bwd");
}

namespace szn {
int embed_argmax_tc(const float* score, const float* table, int n, int D, long long hw, int C, float* scratch,
                    long long* labels, cudaStream_t st);
}

extern "C" long long szn_embed_argmax_scratch_floats(int C, int D) {
  const long long Cpad = C <= 64 ? 64 : C <= 128 ? 128 : 256, Dpad = (D + 31) / 32 * 32;
  const long long tc = 2 * Cpad * Dpad + Cpad;
  return tc > C ? tc : C;
}

// labels[n,h,w] (int64) = argmax_c cos(score[:, :, p], table[c]);  en_scratch: szn_embed_argmax_scratch_floats(C, D) floats
extern "C" int szn_embed_argmax(const float* score, const float* table, int n, int D, int h, int w, int C,
                                float* en_scratch, long long* labels, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int rc = embed_argmax_tc(score, table, n, D, (long long)h * w, C, en_scratch, labels, st);
    if (rc <= 0) return rc;  // launched on the tensor-core path (0) or failed (< 0); 1 = shape not covered, fall through
  }
  table_norm_kernel<<<(C + 127) / 128, 128, 0, st>>>(table, C, D, en_scratch);
  if (int e = check_launch("szn_embed_argmax/norm")) return e;
  const long long hw = (long long)h * w;
  const long long tiles = (hw + 127) / 128 * n;
  embed_argmax_kernel<<<(unsigned)tiles, 256, 0, st>>>(score, table, en_scratch, n, D, hw, C, labels);
  return check_launch("szn_embed_argmax");
}

extern "C" int szn_stitch_labels(const long long* lbl_seen, const long long* lbl_unseen, const float* seen_mask_score,
                                 const long long* target, const long long* unseen, int n_unseen, int n, int h, int w,
                                 long long* out, void* stream) {
  const long long hw = (long long)h * w;
  const int mode = seen_mask_score ? 0 : 1;
  stitch_kernel<<<hgrid(n * hw, 256), 256, 0, (cudaStream_t)stream>>>(lbl_seen, lbl_unseen, seen_mask_score, target, unseen,
                                                                      n_unseen, mode, n, hw, out);
  return check_launch("szn_stitch_labels");
}
