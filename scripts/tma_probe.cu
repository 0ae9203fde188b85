// Development probe (not product code): TMA ingest (L2 -> shared memory) rate of one SM as a function of how many SMs
// pull at the same time, of the box size, of how many boxes are in flight, of whether the CTAs read the same bytes
// (weights) or their own (activations), and with cluster multicast (every CTA of a cluster issues 1/csz of the box and
// all of them receive all of it).  The buffer is L2-resident (32 MB, touched before timing).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/tma_probe scripts/tma_probe.cu -lcuda && build/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct Cfg {
  int iters;       // boxes received per CTA
  int stages;      // boxes in flight per CTA
  int box_rows;    // rows of 128 bytes per box
  int shared_src;  // 1: every CTA (cluster) reads the same region; 0: its own
  int csz;         // cluster size (launch attribute); multicast when > 1
  int region_rows; // rows of the region a CTA (cluster) cycles through
  int planes;      // 1 or 2: the box's third dimension (the hi | lo planes of the convolution kernels' operands)
  int issuers;     // 1 or 2 threads (of different warps), each with its own ring of `stages` boxes
};

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap map, Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[32];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint32_t rank = 0;
  if (c.csz > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const uint32_t box_bytes = (uint32_t)c.box_rows * 128u * (uint32_t)c.planes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < c.stages * c.issuers; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (c.csz > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  const int issuer = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && issuer < c.issuers) {
    const int group = c.csz > 1 ? (int)blockIdx.x / c.csz : (int)blockIdx.x;
    // shared_src: 0 = every CTA (cluster) its own region, 1 = all the same region, g > 1 = groups of g CTAs share one
    const int row0 = c.shared_src == 1 ? 0 : (c.shared_src > 1 ? group / c.shared_src : group) * c.region_rows;
    const int slice_rows = c.box_rows / c.csz;
    const uint16_t mask = (uint16_t)((1u << c.csz) - 1u);
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    uint32_t phase_bits = 0;
    for (int i = 0; i < c.iters + c.stages; ++i) {
      const int s = i % c.stages;
      const uint32_t bar = smem_u32(&full[issuer * c.stages + s]);
      const uint32_t dst = base + (uint32_t)(issuer * c.stages + s) * box_bytes;
      if (i >= c.stages) {
        uint32_t ok = 0;
        const uint32_t par = (phase_bits >> s) & 1u;
        while (!ok)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(ok)
                       : "r"(bar), "r"(par)
                       : "memory");
        phase_bits ^= 1u << s;
      }
      if (i < c.iters) {
        const int r = row0 + ((i * c.issuers + issuer) * c.box_rows) % c.region_rows;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
        if (c.csz == 1) {
          asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
                           "r"(dst), "l"(&map), "r"(bar), "r"(0), "r"(r), "r"(0)
                       : "memory");
        } else {
          asm volatile(
              "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
              "%4, %6}], [%2], %5;" ::"r"(dst + rank * (uint32_t)slice_rows * 128u),
              "l"(&map), "r"(bar), "r"(0), "r"(r + (int)rank * slice_rows), "h"(mask), "r"(0)
              : "memory");
        }
      }
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (issuer == 0) out[2 * blockIdx.x] = (long long)g0;
    __threadfence_block();
    if (issuer == c.issuers - 1) out[2 * blockIdx.x + 1] = (long long)g1;
  }
  __syncthreads();
  if (c.csz > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  const size_t rows_total = 1u << 20;  // x 128 B = 128 MB
  void* buf;
  cudaMalloc(&buf, rows_total * 128);
  cudaMemset(buf, 1, rows_total * 128);
  long long* d_out;
  cudaMalloc(&d_out, sizeof(long long) * 2 * 256);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("sms %d\n", sms);
  printf("%5s %4s %6s %6s %6s %6s %6s | %12s %12s %12s %10s\n", "grid", "csz", "boxKB", "stages", "issuer", "stride", "iters",
         "B/ns per CTA", "min B/ns CTA", "chip GB/s", "ns per box");
  std::vector<long long> out(512);
  struct Case { int grid, csz, box_rows, stages, shared, planes, issuers, row_stride, region_rows; };
  std::vector<Case> cases;
  // The A operand of layer3.conv1 (2048 pixels x 1024 channels, 8 MB, L2-resident): a 128-pixel tile = 16 chunks of
  // 128 rows x 128 B (pitch 2048 B) x 2 planes = 512 KB, walked in 64 KB boxes (2 chunks) by the 8 CTAs that share it
  // (BN = 32), or 4 (BN = 64), or 1.  region_rows counts 128-byte rows per plane of one region.
  for (int gsz : {0, 4, 8, 16})
    for (int stages : {2, 3})
      for (int issuers : {1, 2}) {
        cases.push_back({128, 1, 256, stages, gsz, 2, issuers, 128, 2048});   // 64 KB boxes
        cases.push_back({128, 1, 128, stages, gsz, 2, issuers, 128, 2048});   // 32 KB boxes
      }
  for (int gsz : {0, 8}) cases.push_back({128, 1, 64, 4, gsz, 2, 2, 128, 2048});
  // the same sharing through cluster multicast: each CTA of a cluster issues 1/csz of the box, all receive all of it
  // (the probe has no empty barriers: a CTA that runs a phase ahead of a peer corrupts the peer's transaction count, so
  // only deep rings are safe here)
  if (getenv("PROBE_MULTICAST"))
    for (int csz : {2, 4, 8}) cases.push_back({128, csz, 64, 6, 0, 2, 1, 128, 2048});
  for (const Case& k : cases) {
    if ((size_t)k.stages * k.issuers * k.box_rows * 128 * k.planes + 2048 > (size_t)smem) continue;
    Cfg c{};
    c.iters = 256;
    c.stages = k.stages;
    c.box_rows = k.box_rows;
    c.shared_src = k.shared;
    c.csz = k.csz;
    c.region_rows = k.region_rows;
    c.planes = k.planes;
    c.issuers = k.issuers;
    CUtensorMap map;
    // rows of 128 bytes, `row_stride` bytes apart: the buffer holds rows_total * 128 / row_stride of them per plane pair
    const cuuint64_t rows_avail = rows_total * 128 / (cuuint64_t)k.row_stride / 2;  // per plane
    cuuint64_t dims[3] = {64, rows_avail, 2};
    cuuint64_t strides[2] = {(cuuint64_t)k.row_stride, rows_avail * (cuuint64_t)k.row_stride};
    cuuint32_t box[3] = {64, (cuuint32_t)(k.box_rows / k.csz), (cuuint32_t)k.planes};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("encode failed %d\n", (int)r);
      return 1;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(k.grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = k.csz;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaSuccess;
    for (int it = 0; it < 2 && e == cudaSuccess; ++it) e = cudaLaunchKernelEx(&cfg, probe, map, c, d_out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("grid %d csz %d: %s\n", k.grid, k.csz, cudaGetErrorString(e));
      return 1;
    }
    cudaMemcpy(out.data(), d_out, sizeof(long long) * 2 * k.grid, cudaMemcpyDeviceToHost);
    const double bytes_cta = (double)c.iters * k.box_rows * 128 * k.planes * k.issuers;
    double sum = 0, mn = 1e30;
    long long t0 = out[0], t1 = out[1];
    for (int b = 0; b < k.grid; ++b) {
      const double rate = bytes_cta / (double)(out[2 * b + 1] - out[2 * b]);
      sum += rate;
      mn = std::min(mn, rate);
      t0 = std::min(t0, out[2 * b]);
      t1 = std::max(t1, out[2 * b + 1]);
    }
    printf("%5d %4d %6d %6d %6d %6d %6d | %12.1f %12.1f %12.1f %10.1f  shared by %d\n", k.grid, k.csz, k.box_rows * k.planes / 8, k.stages,
           k.issuers, k.row_stride, c.iters, sum / k.grid, mn, bytes_cta * k.grid / (double)(t1 - t0),
           (double)k.box_rows * 128 * k.planes * k.issuers / (sum / k.grid), k.shared);
    fflush(stdout);
  }
  return 0;
}
