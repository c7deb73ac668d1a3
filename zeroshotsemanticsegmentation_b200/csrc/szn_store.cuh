// Storage formats of trunk activations / data gradients in HBM, as seen by the CUDA-core kernels.
//   float            SZN_F32  : fp32 container holding a TF32-rounded value (operand of kind::tf32 MMAs)
//   __nv_bfloat16    SZN_BF16 : bf16
//   SplitBf16        SZN_F32X3: "fp32-grade" storage for the 3 x bf16 error-compensated products.  A pixel row of C
//                    channels is [hi_0 .. hi_{C-1} | lo_0 .. lo_{C-1}] (2C bf16 = the 4C bytes of an fp32 row) with
//                    hi = bf16(v), lo = bf16(v - hi): v == hi + lo to ~2^-17 relative.  The tensor-core kernel reads the
//                    two planes as separate bf16 operands and issues hi*hi + lo*hi + hi*lo on kind::f16.
// Store<T> moves one 16-byte-per-plane channel vector (VN channels) of a pixel row to / from fp32 registers.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "szn_ptx.cuh"

namespace szn {

struct SplitBf16 {};

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

template <typename T>
struct Store;

template <>
struct Store<float> {
  static constexpr int VN = 4;
  struct Raw {
    uint4 a;
  };
  __host__ __device__ static constexpr size_t row_bytes(long long C) { return (size_t)C * 4; }
  template <bool STREAM>
  __device__ static Raw load_raw(const void* row, int C, int cv) {
    const uint4* p = reinterpret_cast<const uint4*>(row) + cv;
    Raw r;
    r.a = STREAM ? __ldcs(p) : __ldg(p);
    return r;
  }
  template <bool STREAM>
  __device__ static void store_raw(void* row, int C, int cv, const Raw& r) {
    uint4* p = reinterpret_cast<uint4*>(row) + cv;
    if (STREAM) __stcs(p, r.a);
    else *p = r.a;
  }
  __device__ static void to_float(const Raw& r, float (&v)[VN]) {
    v[0] = __uint_as_float(r.a.x), v[1] = __uint_as_float(r.a.y), v[2] = __uint_as_float(r.a.z), v[3] = __uint_as_float(r.a.w);
  }
  __device__ static Raw from_float(const float (&v)[VN]) {
    Raw r;
    r.a = make_uint4(__float_as_uint(to_tf32(v[0])), __float_as_uint(to_tf32(v[1])), __float_as_uint(to_tf32(v[2])),
                     __float_as_uint(to_tf32(v[3])));
    return r;
  }
  __device__ static Raw zero() {
    Raw r;
    r.a = make_uint4(0, 0, 0, 0);
    return r;
  }
  // one element (scalar paths: the 17x17 head gradients)
  __device__ static void store_elem(void* base, long long row, long long ld, int ch, float v) {
    reinterpret_cast<float*>(base)[row * ld + ch] = to_tf32(v);
  }
};

template <>
struct Store<__nv_bfloat16> {
  static constexpr int VN = 8;
  struct Raw {
    uint4 a;
  };
  __host__ __device__ static constexpr size_t row_bytes(long long C) { return (size_t)C * 2; }
  template <bool STREAM>
  __device__ static Raw load_raw(const void* row, int C, int cv) {
    const uint4* p = reinterpret_cast<const uint4*>(row) + cv;
    Raw r;
    r.a = STREAM ? __ldcs(p) : __ldg(p);
    return r;
  }
  template <bool STREAM>
  __device__ static void store_raw(void* row, int C, int cv, const Raw& r) {
    uint4* p = reinterpret_cast<uint4*>(row) + cv;
    if (STREAM) __stcs(p, r.a);
    else *p = r.a;
  }
  __device__ static void to_float(const Raw& r, float (&v)[VN]) {
    const uint32_t w[4] = {r.a.x, r.a.y, r.a.z, r.a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] = __uint_as_float(w[j] << 16);  // bf16 -> fp32 is a 16-bit shift
      v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
    }
  }
  __device__ static Raw from_float(const float (&v)[VN]) {
    Raw r;
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    r.a = make_uint4(w[0], w[1], w[2], w[3]);
    return r;
  }
  __device__ static Raw zero() {
    Raw r;
    r.a = make_uint4(0, 0, 0, 0);
    return r;
  }
  __device__ static void store_elem(void* base, long long row, long long ld, int ch, float v) {
    reinterpret_cast<__nv_bfloat16*>(base)[row * ld + ch] = __float2bfloat16_rn(v);
  }
};

template <>
struct Store<SplitBf16> {
  static constexpr int VN = 8;
  struct Raw {
    uint4 a, b;  // hi, lo
  };
  __host__ __device__ static constexpr size_t row_bytes(long long C) { return (size_t)C * 4; }
  template <bool STREAM>
  __device__ static Raw load_raw(const void* row, int C, int cv) {
    const uint4* p = reinterpret_cast<const uint4*>(row) + cv;
    const uint4* q = p + (C >> 3);  // lo plane: C bf16 = C/8 16-byte vectors further
    Raw r;
    r.a = STREAM ? __ldcs(p) : __ldg(p);
    r.b = STREAM ? __ldcs(q) : __ldg(q);
    return r;
  }
  template <bool STREAM>
  __device__ static void store_raw(void* row, int C, int cv, const Raw& r) {
    uint4* p = reinterpret_cast<uint4*>(row) + cv;
    uint4* q = p + (C >> 3);
    if (STREAM) __stcs(p, r.a), __stcs(q, r.b);
    else *p = r.a, *q = r.b;
  }
  __device__ static void to_float(const Raw& r, float (&v)[VN]) {
    const uint32_t h[4] = {r.a.x, r.a.y, r.a.z, r.a.w}, l[4] = {r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] = __uint_as_float(h[j] << 16) + __uint_as_float(l[j] << 16);
      v[2 * j + 1] = __uint_as_float(h[j] & 0xFFFF0000u) + __uint_as_float(l[j] & 0xFFFF0000u);
    }
  }
  __device__ static Raw from_float(const float (&v)[VN]) {
    Raw r;
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(v[2 * j] - __low2float(hh), v[2 * j + 1] - __high2float(hh));
      h[j] = *reinterpret_cast<const uint32_t*>(&hh);
      l[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    r.a = make_uint4(h[0], h[1], h[2], h[3]);
    r.b = make_uint4(l[0], l[1], l[2], l[3]);
    return r;
  }
  __device__ static Raw zero() {
    Raw r;
    r.a = r.b = make_uint4(0, 0, 0, 0);
    return r;
  }
  __device__ static void store_elem(void* base, long long row, long long ld, int ch, float v) {
    __nv_bfloat16 hi, lo;
    split_bf16(v, hi, lo);
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + row * 2 * ld + ch;
    p[0] = hi;
    p[ld] = lo;
  }
};

// pixel row `row` of a tensor with C channels per row
template <typename T>
__device__ __forceinline__ const void* row_ptr(const void* base, long long row, long long C) {
  return reinterpret_cast<const uint8_t*>(base) + (size_t)row * Store<T>::row_bytes(C);
}
template <typename T>
__device__ __forceinline__ void* row_ptr(void* base, long long row, long long C) {
  return reinterpret_cast<uint8_t*>(base) + (size_t)row * Store<T>::row_bytes(C);
}

}  // namespace szn
