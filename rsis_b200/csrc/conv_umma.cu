// tcgen05 / TMA implicit-GEMM convolution for sm_100a with fp32-grade products from bf16 tensor-core passes.
//
// Replaces the dense contractions of the reference's hot path:
//   * nn.Conv2d + nn.BatchNorm2d (+ReLU, `out += identity`) of /root/reference/src/modules/vision.py:16-19
//     (torchvision Bottleneck.forward) and the skip heads of /root/reference/src/modules/model.py:59-63;
//   * the ConvLSTM gate convolution + sigmoid/tanh + state update of /root/reference/src/modules/clstm.py:43-58 and
//     the global max-pool side feature of model.py:143 (CELL epilogue).
//
// GEMM view: D[M = 128 output pixels][N = BN output channels] += A[M][K] * B[N][K]^T, K = taps x channels.
//   A  activations, split-bf16 planes (hi|lo), NHWC with an arbitrary pixel pitch (so a convolution can read a channel
//      slice of a wider buffer).  Rows of the smem tile are pixels, 64 bf16 channels = 128 bytes each: the K-major
//      SWIZZLE_128B operand layout of tcgen05.mma.  Two ways to stage it:
//        HALO mode (3x3, stride 1, maps at least 8 x 16): one TMA box {64 ch, 10, 18, 1 image, 2 planes} per
//          64-channel chunk brings the 8x16-pixel output tile's whole 10x18 input halo ONCE; the nine filter taps are
//          nine tcgen05 matrix descriptors into that tile (start address shifted by (kh*10+kw) rows, 8-row core
//          matrices 10 rows = 1280 bytes apart), so L2->smem traffic is 1.4x the tile instead of 9x;
//        TAP mode (1x1, stride 2, small maps): one box {64 ch, BW, BH, BI images, 2 planes} per (tap, chunk), start
//          shifted by the tap offset.  Stride-2 convolutions read one of four (row, column)-parity sub-grids per tap.
//      Out-of-bounds pixels/channels are zero-filled by TMA, which is the convolution's zero padding.
//   B  packed weights [2 planes][cout_pad][k_pad] bf16 (rsis_conv_pack_umma), one box {64, BN, 2} per (tap, chunk),
//      in its own smem ring (decoupled from A's: in HALO mode one A stage feeds nine B stages).
//   D  fp32 accumulators in TMEM, two stages of 128 columns so the epilogue of tile i overlaps the MMAs of tile i+1.
//   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi (three kind::f16 MMAs per 16-wide K step, fp32 accumulate): relative
//   product error ~2^-16, which keeps the 104-convolution encoder and the T-step recurrence inside the 1e-3 parity
//   budget (single-pass fp16/bf16/tf32 operands do not: DESIGN.md "precision").
//   K steps that only cover zero padding (channels beyond C in the last chunk) are not issued.
//
// Persistent, warp-specialised CTA (320 threads, one per SM): warps 0-7 epilogue (TMEM lane quarter = warp % 4,
// column half = warp / 4), warp 8 TMA producer, warp 9 MMA issuer; mbarrier full/empty rings for A and B, TMEM
// full/empty between MMA and epilogue.  BN (32/64/128) is chosen so that small maps still spread over the 148 SMs.
// Every mbarrier wait is bounded (trap instead of hang).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace rsis {

extern const bool kHasTcgen05 = true;

namespace {

constexpr int kBM = 128;          // rows (output pixels) per tile == TMEM lanes
constexpr int kBK = 64;           // bf16 channels per K chunk == one 128-byte swizzle row
constexpr int kMaxBN = 128;       // accumulator columns per TMEM stage
constexpr int kAccStages = 2;
constexpr int kTmemCols = kMaxBN * kAccStages;  // 256, power of two
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreadsUmma = kEpiThreads + 64;
constexpr int kMaxStages = 8;
constexpr int kHaloBW = 8, kHaloBH = 16;
constexpr int kHaloRows = (kHaloBW + 2) * (kHaloBH + 2);  // 180 pixels

struct alignas(64) UmmaMaps {
  CUtensorMap a[4];  // stride 1: a[0]; stride 2: one per (row parity, column parity)
  CUtensorMap b;
};

struct UmmaParams {
  // tile geometry
  int BW, BH, BI;          // output-pixel box: width x height x images, product 128
  int tiles_w, tiles_h, tiles_i, tiles_n, num_tiles;
  int BN;                  // output channels per tile (32 / 64 / 128)
  int halo;                // 1: HALO staging, 0: TAP staging
  int a_stages, b_stages;
  int a_plane_bytes;       // bytes of one bf16 plane of an A stage (rows * 128)
  int a_stage_bytes;       // both planes, rounded up to 1024
  int b_stage_bytes;       // 2 * BN * 128
  uint32_t a_tx_bytes, b_tx_bytes;
  uint32_t a_sbo;          // bytes between 8-row core matrices of A
  int base_off_mode;       // debug knob: put (addr >> 7) & 7 into the descriptor's base-offset field
  // K loop
  int taps, ksize, stride, pad;
  int chunks;              // 64-channel chunks of the source
  int last_ksteps;         // 16-wide K steps that carry data in the last chunk (1..4)
  // problem
  int N, Ho, Wo, Cout;
  // conv epilogue
  const float* scale;
  const float* shift;
  View res;
  int res_cs;
  int has_res, relu;
  void* y;
  size_t y_plane;
  int y_fmt, y_cs;
  void* y2;
  size_t y2_plane;
  int y2_fmt, y2_cs;
  // cell epilogue
  const float* c_prev;
  float* h_out;
  float* c_out;
  __nv_bfloat16* h_split;
  size_t hs_plane;
  int hs_cs;
  uint32_t* side_max;
  int side_stride, side_offset;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("rsis_b200 conv_umma: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, cta_group::1
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` when all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row core matrices `sbo` bytes apart.
// The swizzle XOR acts on absolute smem address bits, so a start address shifted by whole rows (HALO taps) or by
// 32 bytes (K step inside the row) addresses the same TMA-written tile.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo, int base_off_mode) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, 16-byte units
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major; canonical 1)
  d |= (uint64_t)(sbo >> 4) << 32;             // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
  if (base_off_mode) d |= (uint64_t)((saddr >> 7) & 7u) << 49;
  d |= (uint64_t)2 << 61;                      // SWIZZLE_128B
  return d;
}

// ---- epilogues --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void store8(void* p, size_t plane, int fmt, size_t idx, const float* v) {
  if (fmt == RSIS_FMT_F32) {
    float4* q = reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + idx);
    q[0] = make_float4(v[0], v[1], v[2], v[3]);
    q[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split_bf16(v[j], hi[j], lo[j]);
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(p);
    *reinterpret_cast<uint4*>(b + idx) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(b + idx + plane) = *reinterpret_cast<const uint4*>(lo);
  }
}
__device__ __forceinline__ void store4u(void* p, size_t plane, int fmt, size_t idx, const float* v) {
  if (fmt == RSIS_FMT_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + idx) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(v[j], hi[j], lo[j]);
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(p);
    *reinterpret_cast<uint2*>(b + idx) = *reinterpret_cast<const uint2*>(hi);
    *reinterpret_cast<uint2*>(b + idx + plane) = *reinterpret_cast<const uint2*>(lo);
  }
}
__device__ __forceinline__ void load4r(const View& v, size_t idx, float* out) {
  if (v.fmt == RSIS_FMT_F32) {
    const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(v.p) + idx);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
  } else {
    const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(v.p);
    const uint2 h = *reinterpret_cast<const uint2*>(b + idx);
    const uint2 l = *reinterpret_cast<const uint2*>(b + idx + v.plane);
    const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(&h);
    const __nv_bfloat16* lp = reinterpret_cast<const __nv_bfloat16*>(&l);
#pragma unroll
    for (int j = 0; j < 4; ++j) out[j] = __bfloat162float(hp[j]) + __bfloat162float(lp[j]);
  }
}

// One 32-column chunk of one accumulator row -> folded BN/bias (+residual) (+ReLU) -> y (and y2).
__device__ __forceinline__ void conv_epilogue_chunk(const UmmaParams& p, const uint32_t (&r)[32], bool row_ok,
                                                    size_t pix, int col0) {
  if (!row_ok) return;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int col = col0 + 4 * q;
    if (col >= p.Cout) break;
    const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + col));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + col));
    float v[4];
    v[0] = fmaf(__uint_as_float(r[4 * q + 0]), sc.x, sh.x);
    v[1] = fmaf(__uint_as_float(r[4 * q + 1]), sc.y, sh.y);
    v[2] = fmaf(__uint_as_float(r[4 * q + 2]), sc.z, sh.z);
    v[3] = fmaf(__uint_as_float(r[4 * q + 3]), sc.w, sh.w);
    if (p.has_res) {
      float t[4];
      load4r(p.res, pix * p.res_cs + col, t);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += t[j];
    }
    if (p.relu) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    store4u(p.y, p.y_plane, p.y_fmt, pix * p.y_cs + col, v);
    if (p.y2) store4u(p.y2, p.y2_plane, p.y2_fmt, pix * p.y2_cs + col, v);
  }
}

// One 32-column chunk = 8 hidden channels x (in, remember, out, cell) of one pixel -> ConvLSTM update (clstm.py:50-58).
// seg = number of consecutive rows (lanes) that belong to the same image (power of two, <= 32) for the side max.
__device__ __forceinline__ void cell_epilogue_chunk(const UmmaParams& p, const uint32_t (&r)[32], bool row_ok,
                                                    size_t pix, int img, int col0, int seg) {
  const int Ch = p.Cout >> 2;
  const int ch0 = col0 >> 2;
  if (ch0 >= Ch) return;  // uniform across the warp
  float hval[8];
  if (row_ok) {
    const size_t idx = pix * Ch + ch0;
    float cp[8];
    if (p.c_prev) {
      const float4 a = *reinterpret_cast<const float4*>(p.c_prev + idx);
      const float4 b = *reinterpret_cast<const float4*>(p.c_prev + idx + 4);
      cp[0] = a.x; cp[1] = a.y; cp[2] = a.z; cp[3] = a.w; cp[4] = b.x; cp[5] = b.y; cp[6] = b.z; cp[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) cp[j] = 0.f;
    }
    float cval[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + col0 + 4 * j));
      const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + col0 + 4 * j));
      const float gi = sigmoidf_acc(fmaf(__uint_as_float(r[4 * j + 0]), sc.x, sh.x));
      const float gf = sigmoidf_acc(fmaf(__uint_as_float(r[4 * j + 1]), sc.y, sh.y));
      const float go = sigmoidf_acc(fmaf(__uint_as_float(r[4 * j + 2]), sc.z, sh.z));
      const float gg = tanhf(fmaf(__uint_as_float(r[4 * j + 3]), sc.w, sh.w));
      const float c = gf * cp[j] + gi * gg;
      cval[j] = c;
      hval[j] = go * tanhf(c);
    }
    store8(p.c_out, 0, RSIS_FMT_F32, idx, cval);
    store8(p.h_out, 0, RSIS_FMT_F32, idx, hval);
    if (p.h_split) store8(p.h_split, p.hs_plane, RSIS_FMT_SPLIT_BF16, pix * p.hs_cs + ch0, hval);
  }
  if (p.side_max) {
    // global nn.MaxPool2d (model.py:143): max over the rows of this warp that belong to the same image, then one
    // atomicMax per (image segment, channel).  Key 0 sorts below every float.
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t key = row_ok ? float_to_key(hval[j]) : 0u;
      for (int s = 1; s < seg; s <<= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, key, s);
        key = o > key ? o : key;
      }
      if ((threadIdx.x & (seg - 1)) == 0 && key != 0u)
        atomicMax(p.side_max + (size_t)img * p.side_stride + p.side_offset + ch0 + j, key);
    }
  }
}

template <bool CELL>
__global__ void __launch_bounds__(kThreadsUmma, 1)
conv_umma_kernel(const __grid_constant__ UmmaMaps maps, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 2 * kAccStages];
  __shared__ uint32_t tmem_slot;

  // SWIZZLE_128B operand tiles need 1024-byte alignment
  const uint32_t smem_a = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_b = smem_a + p.a_stages * p.a_stage_bytes;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t afull0 = smem_u32(&bars[0]);
  const uint32_t aempty0 = smem_u32(&bars[kMaxStages]);
  const uint32_t bfull0 = smem_u32(&bars[2 * kMaxStages]);
  const uint32_t bempty0 = smem_u32(&bars[3 * kMaxStages]);
  const uint32_t tfull0 = smem_u32(&bars[4 * kMaxStages]);
  const uint32_t tempty0 = smem_u32(&bars[4 * kMaxStages + kAccStages]);

  if (warp == kEpiWarps && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    if (p.stride == 2) {
      prefetch_tmap(&maps.a[1]);
      prefetch_tmap(&maps.a[2]);
      prefetch_tmap(&maps.a[3]);
    }
    prefetch_tmap(&maps.b);
  }
  if (warp == kEpiWarps + 1 && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(afull0 + 8 * s, 1);
      mbar_init(aempty0 + 8 * s, 1);
      mbar_init(bfull0 + 8 * s, 1);
      mbar_init(bempty0 + 8 * s, 1);
    }
    for (int a = 0; a < kAccStages; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, kEpiThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  // A "items" per tile: HALO -> one per chunk (nine B items each); TAP -> one per (tap, chunk) (one B item each).
  const int b_per_a = p.halo ? p.taps : 1;
  const int a_items = p.halo ? p.chunks : p.taps * p.chunks;

  if (warp == kEpiWarps) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int nt = tile % p.tiles_n;
        int mt = tile / p.tiles_n;
        const int tw = mt % p.tiles_w;
        mt /= p.tiles_w;
        const int th = mt % p.tiles_h;
        const int ti = mt / p.tiles_h;
        const int w0 = tw * p.BW, h0 = th * p.BH, i0 = ti * p.BI;
        for (int ai = 0; ai < a_items; ++ai) {
          const int cc = p.halo ? ai : ai % p.chunks;
          const int tap0 = p.halo ? 0 : ai / p.chunks;
          mbar_wait(aempty0 + 8 * as, aph ^ 1u);
          {
            const uint32_t sa = smem_a + as * p.a_stage_bytes;
            const uint32_t bar = afull0 + 8 * as;
            mbar_arrive_expect_tx(bar, p.a_tx_bytes);
            if (p.halo) {
              tma_load_5d(sa, &maps.a[0], bar, cc * kBK, w0 - 1, h0 - 1, i0, 0);
            } else {
              const int kh = tap0 / p.ksize, kw = tap0 - kh * p.ksize;
              if (p.stride == 1) {
                tma_load_5d(sa, &maps.a[0], bar, cc * kBK, w0 + kw - p.pad, h0 + kh - p.pad, i0, 0);
              } else {
                // input pixel = 2*out + k - pad: parity (k - pad) & 1, sub-grid index out + floor((k - pad) / 2)
                const int dh = kh - p.pad, dw = kw - p.pad;
                const int ph = dh & 1, pw = dw & 1;
                tma_load_5d(sa, &maps.a[ph * 2 + pw], bar, cc * kBK, w0 + ((dw - pw) >> 1), h0 + ((dh - ph) >> 1), i0,
                            0);
              }
            }
          }
          if (++as == p.a_stages) {
            as = 0;
            aph ^= 1u;
          }
          for (int bi = 0; bi < b_per_a; ++bi) {
            const int tap = tap0 + bi;
            mbar_wait(bempty0 + 8 * bs, bph ^ 1u);
            const uint32_t bar = bfull0 + 8 * bs;
            mbar_arrive_expect_tx(bar, p.b_tx_bytes);
            tma_load_3d(smem_b + bs * p.b_stage_bytes, &maps.b, bar, (tap * p.chunks + cc) * kBK, nt * p.BN, 0);
            if (++bs == p.b_stages) {
              bs = 0;
              bph ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      // kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, N = BN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.BN >> 3) << 17) | ((kBM >> 4) << 24);
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * kMaxBN;
        uint32_t accumulate = 0;
        for (int ai = 0; ai < a_items; ++ai) {
          const int cc = p.halo ? ai : ai % p.chunks;
          const int ksteps = (cc == p.chunks - 1) ? p.last_ksteps : kBK / 16;
          mbar_wait(afull0 + 8 * as, aph);
          tc_fence_after();
          const uint32_t sa = smem_a + as * p.a_stage_bytes;
          for (int bi = 0; bi < b_per_a; ++bi) {
            mbar_wait(bfull0 + 8 * bs, bph);
            tc_fence_after();
            const uint32_t sb = smem_b + bs * p.b_stage_bytes;
            uint32_t sat = sa;
            if (p.halo) {
              const int kh = bi / 3, kw = bi - kh * 3;
              sat += (uint32_t)(kh * (kHaloBW + 2) + kw) * 128u;
            }
            const uint64_t a_hi = make_smem_desc(sat, p.a_sbo, p.base_off_mode);
            const uint64_t a_lo = make_smem_desc(sat + p.a_plane_bytes, p.a_sbo, p.base_off_mode);
            const uint64_t b_hi = make_smem_desc(sb, 1024, 0), b_lo = make_smem_desc(sb + p.BN * 128, 1024, 0);
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t adv = (uint64_t)(k * 32 >> 4);  // 16 bf16 = 32 bytes along K inside the swizzle row
              umma_bf16(d, a_hi + adv, b_hi + adv, idesc, accumulate);
              umma_bf16(d, a_hi + adv, b_lo + adv, idesc, 1u);
              umma_bf16(d, a_lo + adv, b_hi + adv, idesc, 1u);
              accumulate = 1u;
            }
            umma_commit(bempty0 + 8 * bs);  // weight slot free once these MMAs have read it
            if (++bs == p.b_stages) {
              bs = 0;
              bph ^= 1u;
            }
          }
          umma_commit(aempty0 + 8 * as);  // activation slot free
          if (++as == p.a_stages) {
            as = 0;
            aph ^= 1u;
          }
        }
        umma_commit(tfull0 + 8 * acc);  // accumulator complete
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  } else {
    // =============================== epilogue (warps 0-7) ===============================
    const int quarter = warp & 3;       // TMEM lane quarter this warp may read
    const int half = warp >> 2;         // which 32-column chunks (even / odd) this warp handles
    const int row = quarter * 32 + lane;  // accumulator row == TMEM lane
    const int wl = row % p.BW;
    const int hl = (row / p.BW) % p.BH;
    const int il = row / (p.BW * p.BH);
    const int rows_per_img = p.BW * p.BH;
    const int seg = rows_per_img < 32 ? rows_per_img : 32;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nt = tile % p.tiles_n;
      int mt = tile / p.tiles_n;
      const int tw = mt % p.tiles_w;
      mt /= p.tiles_w;
      const int th = mt % p.tiles_h;
      const int ti = mt / p.tiles_h;
      const int wo = tw * p.BW + wl, ho = th * p.BH + hl, img = ti * p.BI + il;
      const bool row_ok = wo < p.Wo && ho < p.Ho && img < p.N;
      const size_t pix = ((size_t)img * p.Ho + ho) * p.Wo + wo;
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kMaxBN;
      for (int c32 = half * 32; c32 < p.BN; c32 += 64) {
        const int col0 = nt * p.BN + c32;
        if (col0 >= p.Cout) break;
        uint32_t r[32];
        tmem_ld32(taddr + c32, r);
        tmem_ld_wait();
        if constexpr (CELL)
          cell_epilogue_chunk(p, r, row_ok, pix, img, col0, seg);
        else
          conv_epilogue_chunk(p, r, row_ok, pix, col0);
      }
      tc_fence_before();
      mbar_arrive(tempty0 + 8 * acc);
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn g_encode = nullptr;
int g_num_sms = 0;
int g_init_status = RSIS_OK;
int g_halo_enabled = 1;    // RSIS_B200_HALO=0 disables HALO staging (debug / A-B timing)
int g_base_off_mode = 0;   // RSIS_B200_HALO_BASEOFF=1 sets the descriptor base-offset field (debug)
int g_min_ctas = 120;      // BN is shrunk until a launch has at least this many tiles (RSIS_B200_MIN_CTAS)
std::once_flag g_once;

constexpr int kSmemLimit = 227 * 1024;
constexpr int kDynSmem = kSmemLimit - 1024;  // static barriers live beside it

void init_once() {
  if (const char* e = getenv("RSIS_B200_HALO")) g_halo_enabled = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_HALO_BASEOFF")) g_base_off_mode = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_MIN_CTAS")) g_min_ctas = atoi(e);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    if (e != cudaSuccess) set_cuda_error(e);
    g_init_status = RSIS_ERR_CUDA;
    return;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  int dev = 0;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess ||
      (e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem)) !=
          cudaSuccess ||
      (e = cudaFuncSetAttribute(conv_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem)) !=
          cudaSuccess) {
    set_cuda_error(e);
    g_init_status = RSIS_ERR_CUDA;
  }
}

int next_pow2(int v) {
  int r = 1;
  while (r < v) r <<= 1;
  return r;
}

// 5-D view {C, W', H', N, plane} of a split-bf16 NHWC activation with pixel pitch `cs` elements
// (sub = 2: one parity sub-grid of a stride-2 conv).
int encode_act_map(CUtensorMap* m, const rsis_tensor& t, int sub, int ph, int pw, int BW, int BH, int BI) {
  const size_t C = t.c, W = t.w, H = t.h, N = t.n, P = pitch(t);
  char* base = reinterpret_cast<char*>(t.data) + ((size_t)ph * W + pw) * P * 2;
  cuuint64_t dims[5] = {C, W / sub, H / sub, N, 2};
  cuuint64_t strides[4] = {P * 2 * sub, W * P * 2 * sub, H * W * P * 2, N * H * W * P * 2};
  cuuint32_t box[5] = {(cuuint32_t)kBK, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BI, 2};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RSIS_OK : RSIS_ERR_CUDA;
}

int encode_weight_map(CUtensorMap* m, const void* w, int cout_pad, int k_pad, int BN) {
  cuuint64_t dims[3] = {(cuuint64_t)k_pad, (cuuint64_t)cout_pad, 2};
  cuuint64_t strides[2] = {(cuuint64_t)k_pad * 2, (cuuint64_t)cout_pad * k_pad * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)BN, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RSIS_OK : RSIS_ERR_CUDA;
}

bool split_ok(const rsis_tensor* t) {
  return valid_tensor(t) && t->fmt == RSIS_FMT_SPLIT_BF16 && aligned16(t->data) && pitch(*t) % 8 == 0;
}
bool out_ok(const rsis_tensor* t) { return valid_tensor(t) && aligned16(t->data) && pitch(*t) % 4 == 0; }

bool common_supported(const rsis_tensor* x, const rsis_conv_weights* w, int stride, int pad) {
  if (!x || !w || !w->w_umma || !w->scale || !w->shift) return false;
  if (w->kh != w->kw || (w->kh != 1 && w->kh != 3) || pad != w->kh / 2) return false;
  if (stride != 1 && stride != 2) return false;
  if (w->cout % 4 != 0 || w->cout < 4) return false;
  if (!split_ok(x) || x->c != w->cin) return false;
  if (stride == 2 && ((x->h & 1) || (x->w & 1))) return false;
  if (!aligned16(w->w_umma) || !aligned16(w->scale) || !aligned16(w->shift)) return false;
  return true;
}

// Fills geometry, tensor maps and the K loop.
int setup(UmmaMaps& maps, UmmaParams& p, const rsis_tensor& x, const rsis_conv_weights* w, int stride, int pad) {
  std::call_once(g_once, init_once);
  if (g_init_status != RSIS_OK) return g_init_status;
  p.N = x.n;
  p.Ho = x.h / stride;
  p.Wo = x.w / stride;
  p.Cout = w->cout;
  p.taps = w->kh * w->kw;
  p.ksize = w->kw;
  p.stride = stride;
  p.pad = pad;
  p.halo = (g_halo_enabled && w->kh == 3 && stride == 1 && p.Wo % kHaloBW == 0 && p.Ho % kHaloBH == 0) ? 1 : 0;
  if (p.halo) {
    p.BW = kHaloBW;
    p.BH = kHaloBH;
    p.BI = 1;
  } else {
    p.BW = next_pow2(p.Wo) < kBM ? next_pow2(p.Wo) : kBM;
    p.BH = next_pow2(p.Ho) < kBM / p.BW ? next_pow2(p.Ho) : kBM / p.BW;
    p.BI = kBM / (p.BW * p.BH);
  }
  p.tiles_w = ceil_div(p.Wo, p.BW);
  p.tiles_h = ceil_div(p.Ho, p.BH);
  p.tiles_i = ceil_div(p.N, p.BI);
  const long long mt = (long long)p.tiles_w * p.tiles_h * p.tiles_i;
  // largest BN that still gives every SM a tile; small maps fall back to narrower tiles
  p.BN = w->cout <= 32 ? 32 : (w->cout <= 64 ? 64 : 128);
  while (p.BN > 32 && mt * ceil_div(w->cout, p.BN) < g_min_ctas) p.BN >>= 1;
  p.tiles_n = ceil_div(w->cout, p.BN);
  const long long nt = mt * p.tiles_n;
  if (nt > 0x7fffffffLL) return RSIS_ERR_UNSUPPORTED;
  p.num_tiles = (int)nt;
  p.chunks = ceil_div(x.c, kBK);
  p.last_ksteps = ceil_div(x.c - (p.chunks - 1) * kBK, 16);
  const int a_rows = p.halo ? kHaloRows : kBM;
  p.a_plane_bytes = a_rows * 128;
  p.a_tx_bytes = (uint32_t)(2 * p.a_plane_bytes);
  p.a_stage_bytes = round_up(2 * p.a_plane_bytes, 1024);
  p.a_sbo = p.halo ? (uint32_t)(kHaloBW + 2) * 128u : 1024u;
  p.base_off_mode = g_base_off_mode;
  p.b_stage_bytes = 2 * p.BN * 128;
  p.b_tx_bytes = (uint32_t)p.b_stage_bytes;
  const int budget = kDynSmem - 1024;
  if (p.halo) {
    p.a_stages = 2;
    p.b_stages = (budget - p.a_stages * p.a_stage_bytes) / p.b_stage_bytes;
    if (p.b_stages > kMaxStages) {
      p.b_stages = kMaxStages;
      p.a_stages = (budget - p.b_stages * p.b_stage_bytes) / p.a_stage_bytes;
      if (p.a_stages > 3) p.a_stages = 3;
    }
  } else {
    p.a_stages = budget / (p.a_stage_bytes + p.b_stage_bytes);
    if (p.a_stages > kMaxStages) p.a_stages = kMaxStages;
    p.b_stages = p.a_stages;
  }
  if (p.a_stages < 1 || p.b_stages < 1) return RSIS_ERR_UNSUPPORTED;
  p.scale = w->scale;
  p.shift = w->shift;
  const int cout_pad = round_up(w->cout, 16);
  const int k_pad = p.taps * p.chunks * kBK;
  if (int e = encode_weight_map(&maps.b, w->w_umma, cout_pad, k_pad, p.BN)) return e;
  if (stride == 1) {
    if (p.halo) {
      if (int e = encode_act_map(&maps.a[0], x, 1, 0, 0, kHaloBW + 2, kHaloBH + 2, 1)) return e;
    } else {
      if (int e = encode_act_map(&maps.a[0], x, 1, 0, 0, p.BW, p.BH, p.BI)) return e;
    }
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        if (int e = encode_act_map(&maps.a[ph * 2 + pw], x, 2, ph, pw, p.BW, p.BH, p.BI)) return e;
  }
  return RSIS_OK;
}

template <bool CELL>
int launch(const UmmaMaps& maps, const UmmaParams& p, cudaStream_t st) {
  const int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
  conv_umma_kernel<CELL><<<grid, kThreadsUmma, kDynSmem, st>>>(maps, p);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // namespace

bool conv2d_umma_supported(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const rsis_tensor* residual,
                           const rsis_tensor* y, const rsis_tensor* y2, int stride, int pad) {
  if (n_src != 1 || !common_supported(srcs, w, stride, pad) || w->gate_interleaved) return false;
  if (!out_ok(y)) return false;
  if (y2 && !out_ok(y2)) return false;
  if (residual && !out_ok(residual)) return false;
  return true;
}

int conv2d_umma(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const rsis_tensor* residual,
                const rsis_tensor* y, const rsis_tensor* y2, int stride, int pad, int relu, cudaStream_t st) {
  (void)n_src;
  UmmaMaps maps;
  UmmaParams p{};
  if (int e = setup(maps, p, srcs[0], w, stride, pad)) return e;
  if (y->n != p.N || y->h != p.Ho || y->w != p.Wo || y->c != p.Cout) return RSIS_ERR_BAD_ARG;
  p.y = y->data;
  p.y_cs = pitch(*y);
  p.y_plane = plane_elems(*y);
  p.y_fmt = y->fmt;
  if (y2) {
    if (y2->n != y->n || y2->h != y->h || y2->w != y->w || y2->c != y->c) return RSIS_ERR_BAD_ARG;
    p.y2 = y2->data;
    p.y2_cs = pitch(*y2);
    p.y2_plane = plane_elems(*y2);
    p.y2_fmt = y2->fmt;
  }
  if (residual) {
    if (residual->n != y->n || residual->h != y->h || residual->w != y->w || residual->c != y->c)
      return RSIS_ERR_BAD_ARG;
    p.res = make_view(*residual);
    p.res_cs = pitch(*residual);
    p.has_res = 1;
  }
  p.relu = relu ? 1 : 0;
  return launch<false>(maps, p, st);
}

bool convlstm_cell_umma_supported(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w) {
  if (n_src != 1 || !w || !w->gate_interleaved || !common_supported(srcs, w, 1, w->kh / 2)) return false;
  return (w->cout / 4) % 8 == 0;  // the epilogue handles 8 hidden channels (32 gate columns) at a time
}

int convlstm_cell_umma(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const float* c_prev,
                       const rsis_tensor* h_out, const rsis_tensor* h_split, const rsis_tensor* c_out,
                       uint32_t* side_max, int side_stride, int side_offset, cudaStream_t st) {
  (void)n_src;
  UmmaMaps maps;
  UmmaParams p{};
  if (int e = setup(maps, p, srcs[0], w, 1, w->kh / 2)) return e;
  const int Ch = p.Cout / 4;
  auto ok = [&](const rsis_tensor* t, int fmt) {
    return valid_tensor(t) && t->fmt == fmt && t->n == p.N && t->h == p.Ho && t->w == p.Wo && t->c == Ch &&
           aligned16(t->data);
  };
  if (!ok(h_out, RSIS_FMT_F32) || !ok(c_out, RSIS_FMT_F32) || pitch(*h_out) != Ch || pitch(*c_out) != Ch)
    return RSIS_ERR_BAD_ARG;
  if (h_split && (!ok(h_split, RSIS_FMT_SPLIT_BF16) || pitch(*h_split) % 8 != 0)) return RSIS_ERR_BAD_ARG;
  if (side_max && (side_stride < side_offset + Ch || side_offset < 0)) return RSIS_ERR_BAD_ARG;
  if (c_prev && !aligned16(c_prev)) return RSIS_ERR_ALIGN;
  p.c_prev = c_prev;
  p.h_out = reinterpret_cast<float*>(h_out->data);
  p.c_out = reinterpret_cast<float*>(c_out->data);
  if (h_split) {
    p.h_split = reinterpret_cast<__nv_bfloat16*>(h_split->data);
    p.hs_cs = pitch(*h_split);
    p.hs_plane = plane_elems(*h_split);
  }
  p.side_max = side_max;
  p.side_stride = side_stride;
  p.side_offset = side_offset;
  return launch<true>(maps, p, st);
}

}  // namespace rsis
