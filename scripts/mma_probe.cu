// Development probe (not product code): issue-to-completion cost of tcgen05.mma kind::f16 on sm_100a as a function of
// the instruction shape (M, N), the A-operand source (SS: shared-memory descriptor, TS: tensor memory) and how many
// accumulators the chain alternates between.  One CTA per SM; one elected thread issues `reps` MMAs back to back,
// commits to an mbarrier and waits.  Prints cycles and ns per MMA (block 0 and the slowest block).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mma_probe scripts/mma_probe.cu && build/mma_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
      "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}

struct Cfg {
  int ts, M, N, nacc, reps, bvar;  // bvar: number of distinct B descriptors cycled through (1..8)
};

__global__ void __launch_bounds__(128, 1) probe(Cfg c, long long* out_cyc, long long* out_ns) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // A: 128 rows x 128 B (16 KB) at base; B: 8 variants x 256 rows x 128 B at base + 16 KB (first 64 KB used twice)
  for (int i = threadIdx.x; i < (16384 + 2 * 32768) / 2; i += blockDim.x)
    reinterpret_cast<__nv_bfloat16*>(gen)[i] = __float2bfloat16(((i * 37) % 17 - 8) * 0.01f);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // zero the TS-mode A region (columns 480..511) so the operand is defined
  {
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 480;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0x3c003c00u)
                 : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr + 8), "r"(0x3c003c00u)
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)c.N >> 3) << 17) | (((uint32_t)c.M >> 4) << 24);
    const uint64_t adesc = make_desc(base, 1024);
    const uint64_t bdesc = make_desc(base + 16384, 1024);
    const int stage_cols = c.N;  // accumulators side by side (nacc * N <= 480)
    long long t0 = clock64();
    unsigned long long g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    for (int r = 0; r < c.reps; ++r) {
      const uint32_t d = tmem + (uint32_t)((r % c.nacc) * stage_cols);
      const int v = r % c.bvar;
      // variants: K advance inside the swizzle row (v & 3) * 32 B, and a second 32 KB tile (v >> 2)
      const uint64_t b = bdesc + (uint64_t)(((v & 3) * 32 + (v >> 2) * 32768) >> 4);
      if (c.ts)
        mma_ts(d, tmem + 480 + (uint32_t)((v & 1) * 8), b, idesc, r >= c.nacc);
      else
        mma_ss(d, adesc + (uint64_t)(((v & 3) * 32) >> 4), b, idesc, r >= c.nacc);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
    long long t_issue = clock64();
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(ok)
          : "r"(smem_u32(&bar))
          : "memory");
    }
    long long t1 = clock64();
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    out_cyc[2 * blockIdx.x] = t1 - t0;
    out_cyc[2 * blockIdx.x + 1] = t_issue - t0;
    out_ns[blockIdx.x] = (long long)(g1 - g0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 16384 + 2 * 32768 + 1024 + 34816;  // + slack: an M=128/N=256 descriptor may read past the tile
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *d_cyc, *d_ns;
  cudaMalloc(&d_cyc, sizeof(long long) * 2 * 256);
  cudaMalloc(&d_ns, sizeof(long long) * 256);
  std::vector<long long> cyc(512), ns(256);
  printf("sms %d\n", sms);
  printf("%-4s %4s %4s %5s %5s %6s %5s | %10s %10s %10s %10s\n", "mode", "M", "N", "nacc", "bvar", "reps", "grid",
         "cyc/mma b0", "cyc/mma mx", "issue cyc", "ns/mma b0");
  const int Ms[2] = {128, 64};
  const int Ns[5] = {256, 128, 64, 32, 16};
  for (int grid : {1, sms}) {
    for (int ts = 0; ts < 2; ++ts)
      for (int M : Ms)
        for (int N : Ns)
          for (int nacc : {1, 2})
            for (int bvar : {1, 8}) {
              if (nacc * N > 480) continue;
              if (grid == sms && (bvar == 1 || (nacc == 2 && N < 128))) continue;
              Cfg c{ts, M, N, nacc, 2048, bvar};
              for (int it = 0; it < 2; ++it) probe<<<grid, 128, smem>>>(c, d_cyc, d_ns);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) {
                printf("%s M=%d N=%d: %s\n", ts ? "TS" : "SS", M, N, cudaGetErrorString(e));
                return 1;
              }
              cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
              cudaMemcpy(ns.data(), d_ns, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
              long long mx = 0;
              for (int b = 0; b < grid; ++b) mx = cyc[2 * b] > mx ? cyc[2 * b] : mx;
              printf("%-4s %4d %4d %5d %5d %6d %5d | %10.1f %10.1f %10.1f %10.1f\n", ts ? "TS" : "SS", M, N, nacc, bvar,
                     c.reps, grid, (double)cyc[0] / c.reps, (double)mx / c.reps, (double)cyc[1] / c.reps,
                     (double)ns[0] / c.reps);
            }
  }
  return 0;
}
