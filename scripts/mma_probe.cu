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

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

struct Cfg {
  int ts, M, N, nacc, reps, bvar;  // bvar: number of distinct B descriptors cycled through (1..8)
  int amode;   // 0: A tile of 8-row groups 1024 B apart, 1024-aligned; 1: HALO staging of the convolution kernels --
               //    groups 1280 B apart (10-pixel rows), start shifted by (kh*10+kw) rows per tap
  int commit;  // > 0: a tcgen05.commit to a scratch mbarrier after every `commit` MMAs
};

template <int TS, int COMMIT>
__global__ void __launch_bounds__(128, 1) probe(Cfg c, long long* out_cyc, long long* out_ns) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // A: 48 KB region at base (aligned tile: 16 KB; halo tile: 180 rows x 128 B); B: 2 x 32 KB at base + 48 KB
  for (int i = threadIdx.x; i < (49152 + 2 * 32768) / 2; i += blockDim.x)
    reinterpret_cast<__nv_bfloat16*>(gen)[i] = __float2bfloat16(((i * 37) % 17 - 8) * 0.01f);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
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
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)c.N >> 3) << 17) | (((uint32_t)c.M >> 4) << 24);
    const uint64_t adesc = make_desc(base, c.amode ? 1280 : 1024);
    const uint64_t bdesc = make_desc(base + 49152, 1024);
    const int stage_cols = c.N;  // accumulators side by side (nacc * N <= 480)
    long long t0 = clock64();
    unsigned long long g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    // tight issue loop: 8 MMAs per iteration, every descriptor precomputed (no integer division on the issue path)
    uint64_t bv[8];
    uint64_t av[8];
    uint32_t atm[8], dv[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const int vv = c.bvar == 1 ? 0 : v;
      bv[v] = bdesc + (uint64_t)(((vv & 3) * 32 + (vv >> 2) * 32768) >> 4);
      av[v] = adesc + (uint64_t)(((vv & 3) * 32 + (c.amode ? ((v % 3) * 10 + (v / 3)) * 128 : 0)) >> 4);
      atm[v] = tmem + 480 + (uint32_t)((vv & 1) * 8);
      dv[v] = tmem + (uint32_t)((v & (c.nacc - 1)) * stage_cols);
    }
    for (int r = 0; r < c.reps; r += 8) {
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        if (TS)
          mma_ts(dv[v], atm[v], bv[v], idesc, 1u);
        else
          mma_ss(dv[v], av[v], bv[v], idesc, 1u);
        if (COMMIT > 0 && (v + 1) % COMMIT == 0)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2))
                       : "memory");
      }
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
  const int smem = 49152 + 2 * 32768 + 1024 + 34816;  // + slack: an M=128/N=256 descriptor may read past the tile
  cudaFuncSetAttribute(probe<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long *d_cyc, *d_ns;
  cudaMalloc(&d_cyc, sizeof(long long) * 2 * 256);
  cudaMalloc(&d_ns, sizeof(long long) * 256);
  std::vector<long long> cyc(512), ns(256);
  printf("sms %d\n", sms);
  printf("%-4s %4s %4s %5s %5s %6s %6s %5s | %10s %10s %10s\n", "mode", "M", "N", "nacc", "amode", "commit", "reps", "grid",
         "cyc/mma b0", "issue cyc", "ns/mma b0");
  for (int grid : {1, sms})
    for (int ts = 0; ts < 2; ++ts)
      for (int N : {256, 128, 64, 32})
        for (int amode = 0; amode < 2; ++amode)
          for (int commit : {0, 4, 1}) {
            if (ts && amode) continue;
            if (grid == sms && commit == 1) continue;
            Cfg c{ts, 128, N, 1, 4096, 8, amode, commit};
            for (int it = 0; it < 2; ++it) {
              if (ts) {
                if (commit == 0) probe<1, 0><<<grid, 128, smem>>>(c, d_cyc, d_ns);
                else if (commit == 4) probe<1, 4><<<grid, 128, smem>>>(c, d_cyc, d_ns);
                else probe<1, 1><<<grid, 128, smem>>>(c, d_cyc, d_ns);
              } else {
                if (commit == 0) probe<0, 0><<<grid, 128, smem>>>(c, d_cyc, d_ns);
                else if (commit == 4) probe<0, 4><<<grid, 128, smem>>>(c, d_cyc, d_ns);
                else probe<0, 1><<<grid, 128, smem>>>(c, d_cyc, d_ns);
              }
            }
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
              printf("%s N=%d amode=%d: %s\n", ts ? "TS" : "SS", N, amode, cudaGetErrorString(e));
              return 1;
            }
            cudaMemcpy(cyc.data(), d_cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
            cudaMemcpy(ns.data(), d_ns, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
            printf("%-4s %4d %4d %5d %5d %6d %6d %5d | %10.1f %10.1f %10.1f\n", ts ? "TS" : "SS", 128, N, 1, amode, commit,
                   c.reps, grid, (double)cyc[0] / c.reps, (double)cyc[1] / c.reps, (double)ns[0] / c.reps);
          }
  return 0;
}
