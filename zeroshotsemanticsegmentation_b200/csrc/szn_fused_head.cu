// EXPERIMENTAL (opt-in, FCN32s(fused_head=True); not on the default path, not yet validated on a GPU):
// cosine loss, nearest-embedding labels and the gradient of the 17x17 score map WITHOUT reading or writing the
// (B, D, H, W) score tensor.  The x32 bilinear upsample (models.py:94,146-147) is linear and per channel, so for an output
// pixel p with taps k (at most 4 nodes of the hs x ws map, weights w_k):
//
//     u_p . e_c   = sum_k w_k (s_k . e_c)                       A = S E^T       (hs*ws x C per image)
//     |u_p|^2     = sum_kl w_k w_l (s_k . s_l)                  G = neighbour Gram entries of S (5 per node)
//     cos_p       = (u_p . e_t) / (|u_p| |e_t|)                 utils.py:75-102
//     label_p     = argmax_c (u_p . e_c) / |e_c|                utils.py:159-185 (|u_p| > 0 is common to all classes)
//     dL/ds_k     = (g/N) [ sum_c M1[k,c] e_c/|e_c| + sum_l M2[k,l] s_l ]
//                   M1[k,c] = sum_{p: t_p = c} w_k(p) a_p,  M2[k,l] = sum_p w_k(p) b_p w_l(p),
//                   a_p = -1/|u_p|,  b_p = cos_p / |u_p|^2
//   kind 1 (mse_loss, utils.py:50-73): |u_p - e_t|^2 = |u_p|^2 - 2 u_p.e_t + |e_t|^2 from the same quantities;
//                   a_p = -2 (on the raw rows e_c), b_p = 2
//
// The algebra is pinned on the CPU by tools/fused_head_math.py + tests/test_fused_head_math.py (float64: loss 1e-10,
// gradient 1e-8, labels exact against the materialised path).  Three small kernels replace four passes over 2.5 GB:
//   fused_nodes_kernel   per node: A row (C dot products of length D) and 5 Gram entries
//   fused_pixels_kernel  per 32x32 pixel block (one tap neighbourhood = "cell"): labels, cos, per-cell M1 / M2 partials
//   fused_grad_kernel    per node: gathers the <= 4 cells that contain it, two tiny contractions -> d s17
#include "szn_internal.h"

namespace szn {

namespace {

struct Ws {
  float *en_inv, *en2, *A, *G, *M1, *M2;
};

__host__ __device__ inline long long ws_floats(int B, int hs, int ws, int C) {
  const long long K = (long long)hs * ws, cells = (long long)(hs + 1) * (ws + 1);
  return 2 * C + B * K * C + B * K * 8 + B * cells * 4 * C + B * cells * 16;
}

inline Ws carve(float* w, int B, int hs, int ws, int C) {
  const long long K = (long long)hs * ws, cells = (long long)(hs + 1) * (ws + 1);
  Ws r;
  r.en_inv = w;
  r.en2 = r.en_inv + C;
  r.A = r.en2 + C;
  r.G = r.A + B * K * C;
  r.M1 = r.G + B * K * 8;
  r.M2 = r.M1 + B * cells * 4 * C;
  return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// en_inv[c] = 1 / |e_c|, 1 for zero rows (utils.py:175); en2[c] = |e_c|^2; one warp per class
__global__ void fused_table_norm_kernel(const float* __restrict__ table, int C, int D, float* __restrict__ en_inv,
                                        float* __restrict__ en2) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  float ss = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = table[(long long)c * D + d];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) en_inv[c] = ss == 0.f ? 1.f : 1.f / sqrtf(ss), en2[c] = ss;
}

// grid (hs*ws, B), block 128, dynamic smem 5*D floats.
// G[(b*K + k)*8 + q]: q = 0 self, 1 right (i,j+1), 2 down (i+1,j), 3 down-right, 4 down-left; 0 when the neighbour is
// outside the map.  A[(b*K + k)*C + c] = s_k . e_c
__global__ void __launch_bounds__(128) fused_nodes_kernel(const float* __restrict__ s17, int ld, int coff,
                                                          const float* __restrict__ table, int D, int hs, int ws, int C,
                                                          float* __restrict__ A, float* __restrict__ G) {
  extern __shared__ float sm[];
  const int k = blockIdx.x, b = blockIdx.y, K = hs * ws;
  const int i = k / ws, j = k - i * ws;
  const float* base = s17 + (long long)b * K * ld + coff;
  const int ni[5] = {i, i, i + 1, i + 1, i + 1}, nj[5] = {j, j + 1, j, j + 1, j - 1};
  for (int q = 0; q < 5; ++q) {
    const bool ok = ni[q] < hs && nj[q] >= 0 && nj[q] < ws;
    const float* src = base + (long long)(ni[q] * ws + nj[q]) * ld;
    for (int d = threadIdx.x; d < D; d += blockDim.x) sm[q * D + d] = ok ? src[d] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int task = warp; task < 5 + C; task += nwarp) {
    const float* v = task < 5 ? sm + task * D : table + (long long)(task - 5) * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc = fmaf(sm[d], v[d], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      if (task < 5) G[((long long)b * K + k) * 8 + task] = acc;
      else A[((long long)b * K + k) * C + (task - 5)] = acc;
    }
  }
}

// grid (ws+1, hs+1, B): one CTA per tap neighbourhood ("cell", anchor node (ia, ja) = (blockIdx.y-1, blockIdx.x-1), taps
// a = 2*da + db at node (ia+da, ja+db)); its pixels are the 32x32 block with y + 19 in [32(ia+1), 32(ia+1)+31].
// block 256: thread = 4 consecutive x of one row.  dynamic smem: Ac[4*C] | einv[C] | e2[C] | M1w[8][4*C] | Gc[16] | M2w[8][16]
template <bool LOSS>
__global__ void __launch_bounds__(256) fused_pixels_kernel(const float* __restrict__ A, const float* __restrict__ G,
                                                           const float* __restrict__ en_inv, const float* __restrict__ en2,
                                                           int kind, const long long* __restrict__ target, int H, int W,
                                                           int hs, int ws, int C, float* __restrict__ M1, float* __restrict__ M2,
                                                           double* __restrict__ accum, long long* __restrict__ labels) {
  extern __shared__ float sm[];
  float* Ac = sm;
  float* einv = Ac + 4 * C;
  float* e2 = einv + C;
  float* M1w = e2 + C;
  float* Gc = M1w + 8 * 4 * C;
  float* M2w = Gc + 16;
  __shared__ double red_a[8], red_b[8];
  const int ja = (int)blockIdx.x - 1, ia = (int)blockIdx.y - 1, b = blockIdx.z;
  const int K = hs * ws, cells = (hs + 1) * (ws + 1), cell = blockIdx.y * (ws + 1) + blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int node[4];
  bool ok[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int ni = ia + (a >> 1), nj = ja + (a & 1);
    ok[a] = ni >= 0 && ni < hs && nj >= 0 && nj < ws;
    node[a] = ok[a] ? ni * ws + nj : 0;
  }
  for (int idx = tid; idx < 4 * C; idx += 256) {
    const int a = idx / C, c = idx - a * C;
    Ac[idx] = ok[a] ? A[((long long)b * K + node[a]) * C + c] : 0.f;
  }
  for (int c = tid; c < C; c += 256) einv[c] = en_inv[c], e2[c] = en2[c];
  if (LOSS)
    for (int idx = tid; idx < 8 * 4 * C; idx += 256) M1w[idx] = 0.f;
  if (tid < 10) {
    // 0..3 self, 4: (0,1) 5: (2,3) 6: (0,2) 7: (1,3) 8: (0,3) 9: (1,2); a missing neighbour already reads as 0 in G
    const int src[10] = {0, 1, 2, 3, 0, 2, 0, 1, 0, 1}, q[10] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 4};
    Gc[tid] = ok[src[tid]] ? G[((long long)b * K + node[src[tid]]) * 8 + q[tid]] : 0.f;
  }
  __syncthreads();

  const int ty = tid >> 3, txb = (tid & 7) * 4;
  const int y = 32 * (ia + 1) - 19 + ty, x0 = 32 * (ja + 1) - 19 + txb;
  const float wy0 = (31.5f - ty) * (1.f / 32.f), wy1 = (ty + 0.5f) * (1.f / 32.f);
  double part = 0.0, cnt = 0.0;
  float m2[10];
#pragma unroll
  for (int e = 0; e < 10; ++e) m2[e] = 0.f;
  if (y >= 0 && y < H) {
#pragma unroll
    for (int qx = 0; qx < 4; ++qx) {
      const int x = x0 + qx, tx = txb + qx;
      if (x < 0 || x >= W) continue;
      const float wx0 = (31.5f - tx) * (1.f / 32.f), wx1 = (tx + 0.5f) * (1.f / 32.f);
      const float w0 = wy0 * wx0, w1 = wy0 * wx1, w2 = wy1 * wx0, w3 = wy1 * wx1;
      float best = -INFINITY;
      int bi = 0;
      for (int c = 0; c < C; ++c) {
        const float pa = fmaf(w0, Ac[c], fmaf(w1, Ac[C + c], fmaf(w2, Ac[2 * C + c], w3 * Ac[3 * C + c])));
        const float v = pa * einv[c];
        if (v > best) best = v, bi = c;  // strict >: the lowest index wins ties
      }
      const long long pix = ((long long)b * H + y) * W + x;
      if (labels) labels[pix] = bi;
      if (LOSS) {
        const long long t = target[pix];
        if (t >= C) part += (double)NAN, cnt += 1.0;  // out-of-range label: poison the loss (like szn_embed_loss_fwd)
        if (t >= 0 && t < C) {
          const float pat = fmaf(w0, Ac[t], fmaf(w1, Ac[C + t], fmaf(w2, Ac[2 * C + t], w3 * Ac[3 * C + t])));
          const float un2 = w0 * w0 * Gc[0] + w1 * w1 * Gc[1] + w2 * w2 * Gc[2] + w3 * w3 * Gc[3] +
                            2.f * (w0 * w1 * Gc[4] + w2 * w3 * Gc[5] + w0 * w2 * Gc[6] + w1 * w3 * Gc[7] + w0 * w3 * Gc[8] +
                                   w1 * w2 * Gc[9]);
          float ap, bp;
          if (kind == 0) {
            const float inv_un = rsqrtf(un2);
            const float cs = pat * inv_un * einv[t];
            part += cs;
            ap = -inv_un, bp = cs * inv_un * inv_un;
          } else {
            part += un2 - 2.f * pat + e2[t];
            ap = -2.f, bp = 2.f;
          }
          cnt += 1.0;
          float* m1 = M1w + warp * 4 * C + (int)t;
          atomicAdd(m1, w0 * ap);
          atomicAdd(m1 + C, w1 * ap);
          atomicAdd(m1 + 2 * C, w2 * ap);
          atomicAdd(m1 + 3 * C, w3 * ap);
          m2[0] = fmaf(w0 * w0, bp, m2[0]), m2[1] = fmaf(w1 * w1, bp, m2[1]);
          m2[2] = fmaf(w2 * w2, bp, m2[2]), m2[3] = fmaf(w3 * w3, bp, m2[3]);
          m2[4] = fmaf(w0 * w1, bp, m2[4]), m2[5] = fmaf(w2 * w3, bp, m2[5]);
          m2[6] = fmaf(w0 * w2, bp, m2[6]), m2[7] = fmaf(w1 * w3, bp, m2[7]);
          m2[8] = fmaf(w0 * w3, bp, m2[8]), m2[9] = fmaf(w1 * w2, bp, m2[9]);
        }
      }
    }
  }
  if constexpr (LOSS) {
#pragma unroll
  for (int e = 0; e < 10; ++e) {
    const float v = warp_sum(m2[e]);
    if (lane == 0) M2w[warp * 16 + e] = v;
  }
  for (int o = 16; o; o >>= 1) {
    part += __shfl_xor_sync(0xffffffffu, part, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) red_a[warp] = part, red_b[warp] = cnt;
  __syncthreads();
  // per-cell partials: plain stores, every cell is owned by exactly one CTA (zeros for cells without valid pixels)
  for (int idx = tid; idx < 4 * C; idx += 256) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += M1w[w * 4 * C + idx];
    M1[((long long)b * cells + cell) * 4 * C + idx] = v;
  }
  if (tid < 16) {
    // symmetric 4x4 from the 10 unique entries
    const int a = tid >> 2, a2 = tid & 3;
    const int lo = a < a2 ? a : a2, hi = a < a2 ? a2 : a;
    int e;
    if (lo == hi) e = lo;
    else if (lo == 0 && hi == 1) e = 4;
    else if (lo == 2 && hi == 3) e = 5;
    else if (lo == 0 && hi == 2) e = 6;
    else if (lo == 1 && hi == 3) e = 7;
    else if (lo == 0 && hi == 3) e = 8;
    else e = 9;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += M2w[w * 16 + e];
    M2[((long long)b * cells + cell) * 16 + tid] = v;
  }
  if (tid == 0) {
    double sa = 0, sb = 0;
    for (int w = 0; w < 8; ++w) sa += red_a[w], sb += red_b[w];
    if (sb != 0.0) {
      atomicAdd(accum, sa);
      atomicAdd(accum + 1, sb);
    }
  }
}  // if constexpr (LOSS)
}

// grid (hs*ws, B), block 128; dynamic smem: m1n[C] | coef[16] ; neighbour node ids in static smem.
// ds[(b*K + k)*ld + ch]: ch in [coff, coff+D) = (gout / N) * gradient, 0 for every other channel of the row.
__global__ void __launch_bounds__(128) fused_grad_kernel(const float* __restrict__ s17, int ld, int coff,
                                                         const float* __restrict__ table, const float* __restrict__ en_inv,
                                                         int kind, int D, int hs, int ws, int C,
                                                         const float* __restrict__ M1, const float* __restrict__ M2,
                                                         const double* __restrict__ accum,
                                                         const float* __restrict__ gout, float* __restrict__ ds) {
  extern __shared__ float sm[];
  float* m1n = sm;
  float* coef = m1n + C;
  __shared__ int nbr[16];
  const int k = blockIdx.x, b = blockIdx.y, K = hs * ws, cells = (hs + 1) * (ws + 1);
  const int i = k / ws, j = k - i * ws;
  const int tid = threadIdx.x;
  // the 4 cells that contain node (i, j): anchors (i - da, j - db); there the node is tap a = 2*da + db
  for (int c = tid; c < C; c += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int ia = i - (a >> 1), ja = j - (a & 1);
      const long long cell = (long long)(ia + 1) * (ws + 1) + (ja + 1);
      v += M1[((long long)b * cells + cell) * 4 * C + a * C + c];
    }
    m1n[c] = kind == 0 ? v * en_inv[c] : v;  // cosine: unit rows e_c/|e_c|; MSE: the rows themselves
  }
  if (tid < 16) {
    const int a = tid >> 2, a2 = tid & 3;
    const int ia = i - (a >> 1), ja = j - (a & 1);
    const long long cell = (long long)(ia + 1) * (ws + 1) + (ja + 1);
    const int ni = ia + (a2 >> 1), nj = ja + (a2 & 1);
    const bool ok = ni >= 0 && ni < hs && nj >= 0 && nj < ws;
    coef[tid] = ok ? M2[((long long)b * cells + cell) * 16 + a * 4 + a2] : 0.f;
    nbr[tid] = ok ? ni * ws + nj : k;
  }
  __syncthreads();
  const float scale = gout[0] / (float)accum[1];
  const float* base = s17 + (long long)b * K * ld + coff;
  float* out = ds + ((long long)b * K + k) * ld;
  for (int ch = tid; ch < ld; ch += blockDim.x) {
    const int d = ch - coff;
    float acc = 0.f;
    if (d >= 0 && d < D) {
      for (int c = 0; c < C; ++c) acc = fmaf(m1n[c], table[(long long)c * D + d], acc);
#pragma unroll
      for (int q = 0; q < 16; ++q) acc = fmaf(coef[q], base[(long long)nbr[q] * ld + d], acc);
      acc *= scale;
    }
    out[ch] = acc;
  }
}

}  // namespace
}  // namespace szn
using namespace szn;

extern "C" long long szn_head_fused_workspace_floats(int B, int hs, int ws, int C) { return ws_floats(B, hs, ws, C); }

static int fused_common(const float* s17, int ld, int coff, const float* table, int B, int D, int hs, int ws, int C,
                        float* workspace, Ws* out, cudaStream_t st) {
  if (B < 1 || D < 1 || C < 1 || hs < 1 || ws < 1 || ld < coff + D)
    return set_error(SZN_ERR_ARG, "szn_head_fused: bad shape");
  if ((size_t)(8 * 4 * C + 5 * C + 160) * 4 > 200 * 1024 || (size_t)5 * D * 4 > 200 * 1024)
    return set_error(SZN_ERR_UNSUPPORTED, "szn_head_fused: C or D too large for the shared-memory tables");
  *out = carve(workspace, B, hs, ws, C);
  fused_table_norm_kernel<<<(C + 3) / 4, 128, 0, st>>>(table, C, D, out->en_inv, out->en2);
  if (int e = check_launch("szn_head_fused/norm")) return e;
  const size_t smem = (size_t)5 * D * sizeof(float);
  static size_t nodes_smem = 48 * 1024;
  if (smem > nodes_smem) {
    if (cudaFuncSetAttribute(fused_nodes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return set_error(SZN_ERR_CUDA, "szn_head_fused: shared memory attribute");
    nodes_smem = smem;
  }
  fused_nodes_kernel<<<dim3(hs * ws, B), 128, smem, st>>>(s17, ld, coff, table, D, hs, ws, C, out->A, out->G);
  return check_launch("szn_head_fused/nodes");
}

static size_t pixels_smem(int C) { return (size_t)(4 * C + 2 * C + 8 * 4 * C + 16 + 8 * 16) * sizeof(float); }

template <bool LOSS>
static int launch_pixels(const Ws& w, int kind, const long long* target, int B, int H, int W, int hs, int ws, int C,
                         double* accum, long long* labels, cudaStream_t st) {
  const size_t smem = pixels_smem(C);
  static size_t cur = 48 * 1024;
  if (smem > cur) {
    if (cudaFuncSetAttribute(fused_pixels_kernel<LOSS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return set_error(SZN_ERR_CUDA, "szn_head_fused: shared memory attribute");
    cur = smem;
  }
  fused_pixels_kernel<LOSS><<<dim3(ws + 1, hs + 1, B), 256, smem, st>>>(w.A, w.G, w.en_inv, w.en2, kind, target, H, W, hs, ws,
                                                                       C, w.M1, w.M2, accum, labels);
  return check_launch("szn_head_fused/pixels");
}

extern "C" int szn_head_fused_fwd(int kind, const float* s17, int ld, int coff, const long long* target, const float* table,
                                  int B, int D, int H, int W, int hs, int ws, int C, float* workspace, double* accum,
                                  float* loss, long long* labels, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  // every output pixel must find its taps inside the (hs+1) x (ws+1) cell grid: y + 19 < 32 * (hs + 1)
  if (H + 19 > 32 * (hs + 1) || W + 19 > 32 * (ws + 1)) return set_error(SZN_ERR_ARG, "szn_head_fused_fwd: H/W vs hs/ws");
  if (kind != 0 && kind != 1) return set_error(SZN_ERR_ARG, "szn_head_fused_fwd: kind");
  Ws w{};
  if (int e = fused_common(s17, ld, coff, table, B, D, hs, ws, C, workspace, &w, st)) return e;
  if (target) {
    if (!accum || !loss) return set_error(SZN_ERR_ARG, "szn_head_fused_fwd: accum / loss");
    cudaMemsetAsync(accum, 0, 2 * sizeof(double), st);
    if (int e = launch_pixels<true>(w, kind, target, B, H, W, hs, ws, C, accum, labels, st)) return e;
    return szn_loss_finalize(kind, accum, loss, stream);
  }
  if (!labels) return set_error(SZN_ERR_ARG, "szn_head_fused_fwd: nothing to compute");
  return launch_pixels<false>(w, kind, nullptr, B, H, W, hs, ws, C, nullptr, labels, st);
}

extern "C" int szn_head_fused_bwd(int kind, const float* s17, int ld, int coff, const float* table, int B, int D, int hs,
                                  int ws, int C, const float* workspace, const double* accum, const float* grad_out, float* ds17,
                                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (B < 1 || D < 1 || C < 1 || ld < coff + D) return set_error(SZN_ERR_ARG, "szn_head_fused_bwd: bad shape");
  const Ws w = carve(const_cast<float*>(workspace), B, hs, ws, C);
  const size_t smem = (size_t)(C + 16) * sizeof(float);
  fused_grad_kernel<<<dim3(hs * ws, B), 128, smem, st>>>(s17, ld, coff, table, w.en_inv, kind, D, hs, ws, C, w.M1, w.M2, accum,
                                                         grad_out, ds17);
  return check_launch("szn_head_fused_bwd");
}
