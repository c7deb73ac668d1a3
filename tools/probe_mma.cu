// Micro-benchmark (round-2 tool, NOT part of libszn.so): cycles per tcgen05.mma as a function of M, N, kind and the
// number of TMEM accumulators the K-steps rotate over.  It answers the question DESIGN.md §4.1 leaves open for the
// narrow layers (conv1_2: N = 64, conv2_x: N = 128): is a ~130-cycle floor per instruction real, and would M = 64 x N = 256
// (weights as the A operand, pixels as B) do twice the work per instruction?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe_mma tools/probe_mma.cu
//   ./tools/probe_mma            (on the GPU box; prints one line per configuration)
//
// One CTA per SM (148), one thread issues `iters` MMAs (K-major SWIZZLE_128B operands resident in shared memory, zero
// filled: the tensor pipe does not care about values), rotating over `nacc` accumulators, then commits and waits.
// cycles/MMA = (clock64 after the commit's mbarrier flips - clock64 before the first issue) / iters.
#include <cstdio>
#include <cstdlib>

#include "../zeroshotsemanticsegmentation_b200/csrc/szn_ptx.cuh"

using namespace szn;

struct Cfg {
  int tf32, M, N, nacc, iters, issue;
};

// A operand read from TENSOR MEMORY (".ts" form: tcgen05.mma [d], [a_tmem], b_desc, ...): does an instruction still pay the
// ~32 cycles a 128 x 32-byte A tile costs when it comes from shared memory?
template <bool kTF32>
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (kTF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Cfg c, long long* cycles_out) {
  extern __shared__ uint8_t raw[];
  const uint32_t a0 = smem_u32(raw);
  uint8_t* smem = raw + ((1024u - (a0 & 1023u)) & 1023u);
  // A: up to 128 rows x 128 B, B: up to 256 rows x 128 B, 4 stages each so that consecutive MMAs read other addresses
  constexpr int A_BYTES = 128 * 128, B_BYTES = 256 * 128, STAGES = 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < STAGES * (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(tptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  // issue == 0: `if (threadIdx.x == 0)` (round-1 style: ptxas wraps every UTCHMMA in an ELECT/BRA.U.ANY loop)
  // issue == 1: whole warp in the branch, one elected lane issues (what the kernels do now)
  const uint32_t idesc = umma_idesc(c.tf32 ? 2 : 1, 0, 0, c.M, c.N);
  const uint32_t a_base = smem_u32(smem), b_base = a_base + STAGES * A_BYTES;
  const int acc_cols = c.N < 32 ? 32 : c.N;
  if (c.issue == 0) {
    if (threadIdx.x == 0) {
      const long long t0 = clock64();
      for (int it = 0; it < c.iters; ++it) {
        const int s = it & (STAGES - 1), k = (it >> 2) & 3, acc = it & (c.nacc - 1);
        const uint64_t ad = umma_desc_sw128(a_base + s * A_BYTES + k * 32, 16, 1024);
        const uint64_t bd = umma_desc_sw128(b_base + s * B_BYTES + k * 32, 16, 1024);
        if (c.tf32) tc_mma<true>(tmem + (uint32_t)(acc * acc_cols), ad, bd, idesc, (uint32_t)(it >= c.nacc));
        else tc_mma<false>(tmem + (uint32_t)(acc * acc_cols), ad, bd, idesc, (uint32_t)(it >= c.nacc));
      }
      tc_commit(bar);
      mbar_wait(bar, 0);
      cycles_out[blockIdx.x] = clock64() - t0;
    }
  } else if (warp_idx() == 0) {
    const long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < c.iters; ++it) {
        const int s = it & (STAGES - 1), k = (it >> 2) & 3, acc = it & (c.nacc - 1);
        const uint64_t ad = umma_desc_sw128(a_base + s * A_BYTES + k * 32, 16, 1024);
        const uint64_t bd = umma_desc_sw128(b_base + s * B_BYTES + k * 32, 16, 1024);
        if (c.issue == 3) {
          // A: M lanes x 8 columns (K = 8 tf32 or 16 packed bf16) of uninitialised tensor memory above the accumulators
          // (columns 384..511, 16 slots taken in turn); values do not matter to the tensor pipe's timing
          const uint32_t at = tmem + 384u + (uint32_t)((it & 15) * 8);
          if (c.tf32) tc_mma_ts<true>(tmem + (uint32_t)(acc * acc_cols), at, bd, idesc, (uint32_t)(it >= c.nacc));
          else tc_mma_ts<false>(tmem + (uint32_t)(acc * acc_cols), at, bd, idesc, (uint32_t)(it >= c.nacc));
        } else if (c.tf32) tc_mma<true>(tmem + (uint32_t)(acc * acc_cols), ad, bd, idesc, (uint32_t)(it >= c.nacc));
        else tc_mma<false>(tmem + (uint32_t)(acc * acc_cols), ad, bd, idesc, (uint32_t)(it >= c.nacc));
      }
      tc_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

// CTA pair (cta_group::2): M = 256 over two SMs, each SM holds 128 rows of A and N/2 rows of B.  cycles per pair-MMA as seen
// by the leader; per SM the instruction does 128 x N x K MACs, like the single-CTA M = 128 instruction.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2_kernel(Cfg c, long long* cycles_out) {
  extern __shared__ uint8_t raw[];
  const uint32_t a0 = smem_u32(raw);
  uint8_t* smem = raw + ((1024u - (a0 & 1023u)) & 1023u);
  constexpr int A_BYTES = 128 * 128, B_BYTES = 256 * 128, STAGES = 4;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < STAGES * (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp_idx() == 0) {
    tmem_alloc2(tptr, 512);
    tmem_relinquish2();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tptr;
  const uint32_t idesc = umma_idesc(c.tf32 ? 2 : 1, 0, 0, 256, c.N);
  const uint32_t a_base = smem_u32(smem), b_base = a_base + STAGES * A_BYTES;
  const int acc_cols = c.N < 32 ? 32 : c.N;
  if (rank == 0 && warp_idx() == 0) {
    const long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < c.iters; ++it) {
        const int s = it & (STAGES - 1), k = (it >> 2) & 3, acc = it & (c.nacc - 1);
        const uint64_t ad = umma_desc_sw128(a_base + s * A_BYTES + k * 32, 16, 1024);
        const uint64_t bd = umma_desc_sw128(b_base + s * B_BYTES + k * 32, 16, 1024);
        if (c.tf32) tc_mma2<true>(tmem + (uint32_t)(acc * acc_cols), ad, bd, idesc, (uint32_t)(it >= c.nacc));
        else tc_mma2<false>(tmem + (uint32_t)(acc * acc_cols), ad, bd, idesc, (uint32_t)(it >= c.nacc));
      }
      tc_commit2(bar, 1);  // the leader's barrier only
    }
    __syncwarp();
    mbar_wait(bar, 0);
    if (threadIdx.x == 0) cycles_out[blockIdx.x] = clock64() - t0;
  } else if (threadIdx.x == 0) {
    cycles_out[blockIdx.x] = 0;
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp_idx() == 0) tmem_dealloc2(tmem, 512);
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const size_t smem = 4 * (128 * 128 + 256 * 128) + 1024 + 64;
  if (cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    printf("cudaFuncSetAttribute failed\n");
    return 1;
  }
  long long* d = nullptr;
  cudaMalloc(&d, sizeof(long long) * sms);
  long long* h = (long long*)malloc(sizeof(long long) * sms);
  printf("device %d: %d SMs, nominal %.0f MHz\n", dev, sms, khz / 1000.0);
  printf("%-5s %4s %4s %5s | %10s | %10s | %s\n", "kind", "M", "N", "nacc", "cyc/MMA", "MAC/cyc/SM", "of the 1x-rate floor max(M,128)*N/256");
  const int Ms[2] = {128, 64}, Ns[5] = {32, 64, 128, 192, 256}, accs[3] = {1, 2, 4};
  for (int issue = 0; issue < 2; ++issue)
  for (int tf32 = 1; tf32 >= 0; --tf32)
    for (int mi = 0; mi < (issue ? 2 : 1); ++mi)
      for (int ni = 0; ni < 5; ++ni)
        for (int ai = 0; ai < 3; ++ai) {
          if (!issue && (ai || (ni != 1 && ni != 4))) continue;  // the old issue style: two reference lines only
          Cfg c{tf32, Ms[mi], Ns[ni], accs[ai], 4096, issue};
          const int acc_cols = c.N < 32 ? 32 : c.N;
          if (c.nacc * acc_cols > 512) continue;
          for (int rep = 0; rep < 2; ++rep) {  // first launch warms up
            probe_kernel<<<sms, 128, smem>>>(c, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("launch failed (%s) for kind=%s M=%d N=%d nacc=%d\n", cudaGetErrorString(e), tf32 ? "tf32" : "bf16", c.M,
                     c.N, c.nacc);
              return 1;
            }
          }
          cudaMemcpy(h, d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
          long long worst = 0;
          for (int i = 0; i < sms; ++i) worst = h[i] > worst ? h[i] : worst;
          const double cyc = (double)worst / c.iters;
          const int kk = tf32 ? 8 : 16;
          const double floor_c = (double)(c.M > 128 ? c.M : 128) * c.N / 256.0;
          printf("%-5s %4d %4d %5d | %10.1f | %10.0f | %.2fx%s\n", tf32 ? "tf32" : "bf16", c.M, c.N, c.nacc, cyc,
                 (double)c.M * c.N * kk / cyc, cyc / floor_c, issue ? "" : "   [threadIdx.x == 0 issue]");
        }
  // ---- A from tensor memory ----
  if (getenv("PROBE_TS")) {
    printf("A operand in tensor memory (.ts), M = 128, elected-lane issue\n");
    for (int tf32 = 1; tf32 >= 0; --tf32)
      for (int ni = 0; ni < 5; ++ni) {
        Cfg c{tf32, 128, Ns[ni], 1, 4096, 3};
        for (int rep = 0; rep < 2; ++rep) {
          probe_kernel<<<sms, 128, smem>>>(c, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("ts launch failed (%s) for kind=%s N=%d\n", cudaGetErrorString(e), tf32 ? "tf32" : "bf16", c.N);
            return 1;
          }
        }
        cudaMemcpy(h, d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        long long worst = 0;
        for (int i = 0; i < sms; ++i) worst = h[i] > worst ? h[i] : worst;
        const double cyc = (double)worst / c.iters;
        printf("%-5s %4d %4d %5d | %10.1f | %10.0f | A in TMEM\n", tf32 ? "tf32" : "bf16", c.M, c.N, c.nacc, cyc,
               (double)c.M * c.N * (tf32 ? 8 : 16) / cyc);
      }
    cudaFree(d);
    free(h);
    return 0;
  }
  // ---- CTA pairs ----
  if (cudaFuncSetAttribute(probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    printf("cudaFuncSetAttribute (pair) failed\n");
    return 1;
  }
  printf("cta_group::2 (M = 256 over a CTA pair; MAC/cyc/SM counts 128 x N x K per SM)\n");
  for (int tf32 = 1; tf32 >= 0; --tf32)
    for (int ni = 1; ni < 5; ++ni)
      for (int ai = 0; ai < 2; ++ai) {
        Cfg c{tf32, 256, Ns[ni], accs[ai], 4096, 2};
        const int acc_cols = c.N < 32 ? 32 : c.N;
        if (c.nacc * acc_cols > 512) continue;
        for (int rep = 0; rep < 2; ++rep) {
          probe2_kernel<<<sms & ~1, 128, smem>>>(c, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("pair launch failed (%s) for kind=%s N=%d nacc=%d\n", cudaGetErrorString(e), tf32 ? "tf32" : "bf16", c.N, c.nacc);
            return 1;
          }
        }
        cudaMemcpy(h, d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        long long worst = 0;
        for (int i = 0; i < (sms & ~1); ++i) worst = h[i] > worst ? h[i] : worst;
        const double cyc = (double)worst / c.iters;
        const int kk = tf32 ? 8 : 16;
        printf("%-5s %4d %4d %5d | %10.1f | %10.0f | %.2fx of N/2\n", tf32 ? "tf32" : "bf16", c.M, c.N, c.nacc, cyc,
               128.0 * c.N * kk / cyc, cyc / (c.N / 2.0));
      }
  cudaFree(d);
  free(h);
  return 0;
}
