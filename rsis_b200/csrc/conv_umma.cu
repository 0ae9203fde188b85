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
//   a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi (fp32 accumulate; for BN <= 128 as TWO MMAs per 16-wide K step: X_hi x [W_hi | W_lo]
//   with N = 2 BN and X_lo x W_hi with N = BN; three MMAs of N = BN for BN = 256): relative
//   product error ~2^-16, which keeps the 104-convolution encoder and the T-step recurrence inside the 1e-3 parity
//   budget (single-pass fp16/bf16/tf32 operands do not: DESIGN.md "precision").
//   K steps that only cover zero padding (channels beyond C in the last chunk) are not issued.
//
// Persistent, warp-specialised CTA (352 threads, one per SM): warps 0-7 epilogue (TMEM lane quarter = warp % 4,
// column half = warp / 4), warp 8 activation-TMA producer, warp 9 MMA issuer, warp 10 weight-TMA producer; mbarrier
// full/empty rings for A and B, TMEM full/empty between MMA and epilogue.  A host-side planner picks the
// output-channel width BN (32..256) and a K split per launch.  Every mbarrier wait is bounded (trap, not hang).
#include <cuda.h>

#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace rsis {

extern const bool kHasTcgen05 = true;

namespace {

constexpr int kBM = 128;          // rows (output pixels) per tile == TMEM lanes
constexpr int kBK = 64;           // bf16 channels per K chunk == one 128-byte swizzle row
constexpr int kMaxBN = 256;       // widest output-channel tile
constexpr int kAccStages = 2;
constexpr int kStageCols = 256;   // accumulator columns per TMEM stage (BN, or 2*BN when the weight planes are stacked)
constexpr int kTmemCols = kStageCols * kAccStages;  // 512: the whole TMEM (one CTA per SM)
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreadsUmma = kEpiThreads + 96;  // + A producer, MMA issuer, B producer
constexpr int kMaxStages = 9;  // 9 = all taps of a one-chunk 3x3 weight set resident (b_resident)
constexpr int kHaloBW = 8, kHaloBH = 16;
constexpr int kHaloRows = (kHaloBW + 2) * (kHaloBH + 2);  // 180 pixels

struct alignas(64) UmmaMaps {
  CUtensorMap a[4];  // stride 1: a[0]; stride 2: one per (row parity, column parity)
  CUtensorMap b;
  CUtensorMap pre;   // cells with pre_tma: the hoisted gate share [N][H][W][4*Ch] fp32, boxes of 32 columns
};

// Division by a launch constant without the ~30-instruction integer-division sequence: q = umulhi(n, ceil(2^32 / d)),
// exact for n * d < 2^32 (tile counts are < 2^22 and the divisors used here < 2^10; larger divisors keep magic = 0 and
// divide for real).  The per-tile paths of every warp role decode (tile -> n tile, w, h, image) with these: with
// runtime divisions the decode alone cost a few microseconds per tile on the latency-bound single-warp roles.
struct FastDiv {
  uint32_t d, magic;
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = (uint32_t)d;
  f.magic = (d > 1 && d < 1024) ? (uint32_t)((0x100000000ULL + (uint32_t)d - 1) / (uint32_t)d) : 0u;
  return f;
}
__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f) {
  if (f.d == 1) return n;
  return f.magic ? __umulhi(n, f.magic) : n / f.d;
}

struct UmmaParams {
  // tile geometry
  int BW, BH, BI;          // output-pixel box: width x height x images, product 128
  int tiles_w, tiles_h, tiles_i, tiles_n, num_tiles;
  FastDiv dn, dw, dh, dks; // dividers by tiles_n, tiles_w, tiles_h, ksplit
  int BN;                  // output channels per tile (32 / 64 / 128 / 256)
  int stacked;             // 1: one MMA multiplies by [W_hi | W_lo] (N = 2*BN), 2 MMAs per K step; 0: 3 MMAs, N = BN
  int pw;                  // epilogue piece width in accumulator columns (32, or 16 so that BN = 32 keeps all 8 warps busy)
  int ksplit;              // K slices per tile (split-K across co-resident CTAs, reduced through `scratch`)
  float* scratch;          // [num_tiles * ksplit][128][ncols] fp32 partial accumulators
  unsigned* counters;      // [num_tiles][2] arrival / completion counters (zero between launches)
  int halo;                // 1: HALO staging, 0: TAP staging
  int a_stages, b_stages;
  int b_resident;          // 1: the tile's whole weight set stays in the B ring for the CTA's lifetime (loaded once)
  int a_plane_bytes;       // bytes of one bf16 plane of one 64-channel chunk of an A stage (rows * 128)
  int a_chunk_bytes;       // both planes of one chunk, rounded up to 1024
  int a_stage_bytes;       // ag * a_chunk_bytes = one TMA box
  int b_chunk_bytes;       // planes * BN * 128: the weights of one (tap, chunk) K block
  int b_stage_bytes;       // bg * b_chunk_bytes = one TMA box
  // One elected thread issues a TMA box every ~230-400 ns whatever the box holds (scripts/tma_probe.cu, in-kernel
  // stamps: profiles/r2v_*), so 8 KB weight boxes starve 25 ns MMAs.  Boxes therefore carry several K blocks:
  int ag;                  // 64-channel chunks per activation box (flat-pixel 1x1 convolutions: up to 2); otherwise 1
  int bg;                  // K blocks per weight box: HALO -> taps of one chunk (1 / 3 / 9), TAP -> chunks of one tap
  int a_flat;              // 1: maps.a[0] is the 4-D flat-pixel view {64, pixels, plane, chunk} of a 1x1 stride-1 input
  int csplit;              // 1: the two K slices of a tile (ksplit == 2) are the two CTAs of a CLUSTER; rank 1 hands its partial
                           //    accumulator to rank 0 through distributed shared memory (no global scratch, no counters)
  int x_off;               // csplit: byte offset of the exchange buffer [128 rows][x_pitch floats] in dynamic smem (0: it
                           //    reuses rank 0's operand stages, which are dead once rank 0's accumulator is complete)
  int x_pitch;             // csplit: floats per exchange row (BN + 4: conflict-free 16-byte accesses with thread = row)
  int early_b;             // 1: the weight producer does not wait for the previous kernel of the stream (static weights)
  int a_split_off;         // a_sw64 == 2: byte offset of the 16-channel SWIZZLE_32B part inside an activation stage
  int a_sw64;              // 2: 33..48-channel cell sources as TWO boxes per tile, 32 channels SWIZZLE_64B + 16 channels
                           //    SWIZZLE_32B (34.5 KB per stage instead of 46 KB), so that the nine taps of [W_hi | W_lo]
                           //    (144 KB at 64 gate columns) stay RESIDENT beside two stages instead of streaming per tile;
                           // 1: HALO boxes of 32 channels in SWIZZLE_64B rows of 64 bytes (sources of <= 32 channels: the TMA
                           //    unit spends ~1.6 ns per box ROW and twice that on rows that are mostly zero fill --
                           //    scripts/halo_probe.cu, profiles/r2ba_halo_probe.txt: 1217 -> 564 ns per 24-channel halo box)
  int b_map3d;             // 1: maps.b is the 3-D view {k, cout, plane} (one K block per box)
  int agroups;             // activation boxes per tap: HALO 1 (the walk is chunk-major), TAP ceil(chunks / ag)
  uint32_t a_tx_bytes, b_tx_bytes;
  uint32_t a_sbo;          // bytes between 8-row core matrices of A
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
  const float* preact;     // optional [N][H][W][4*Ch] fp32, (channel, gate) order: time-invariant part of the gates
  float* h_out;
  float* c_out;
  __nv_bfloat16* h_split;
  size_t hs_plane;
  int hs_cs;
  uint32_t* side_max;
  int side_stride, side_offset;
  int cell_rows;           // 1: row-wise cell epilogue (thread = pixel, no shared-memory transpose); 0: staged transpose
  int single;              // 1: single-pass bf16 (training mode of BASELINE.json configs[3]): only the hi planes are loaded
                           //    and multiplied -- one MMA of N = BN per K step instead of 2 (stacked) or 3
  int pre_tma;             // 1: the hoisted gate share of a tile is staged in shared memory by TMA (two stages)
  int p_stage_bytes;       // BN/32 boxes of 128 rows x 128 bytes
#ifdef RSIS_DEBUG_TIMING
  unsigned long long* trace;  // per-launch trace row (24 x u64) of rsis_debug_trace, or nullptr
  int dbg_skip;               // RSIS_B200_DBG_SKIP: 1 = the cell epilogue loads no state / gate share, 2 = stores nothing
#endif
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

// ---- cluster (distributed shared memory) helpers for the cluster split-K
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {  // acquire at cluster scope, bounded
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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

// One lane of a fully converged warp; the compiler keeps the elected region's operands in uniform registers, so
// UTMALDG / UTCHMMA issue back to back (an `if (lane == 0)` region makes it wrap every one in a uniformisation loop).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
template <int PW>
__device__ __forceinline__ void tmem_ld_piece(uint32_t taddr, uint32_t (&r)[PW]) {
  if constexpr (PW == 32)
    tmem_ld32(taddr, r);
  else
    tmem_ld16(taddr, r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 128 bytes, 8-row core matrices `sbo` bytes apart.
// The swizzle XOR acts on absolute smem address bits, so a start address shifted by whole rows (HALO taps) or by
// 32 bytes (K step inside the row) addresses the same TMA-written tile.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t sbo, uint32_t layout = 2u) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address, 16-byte units
  d |= (uint64_t)1 << 16;                      // leading byte offset (unused for swizzled K-major; canonical 1)
  d |= (uint64_t)(sbo >> 4) << 32;             // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                      // descriptor version (sm_100)
  d |= (uint64_t)layout << 61;                 // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B (rows of 64 bytes, 512-byte atoms)
  return d;
}

// ---- epilogues --------------------------------------------------------------------------------------------------
// The accumulator comes out of TMEM with thread = row (pixel) and registers = columns (channels), but NHWC memory
// wants consecutive lanes on consecutive channels.  Every 32-row x PW-column piece is therefore transposed through a
// per-warp shared-memory tile (pitch 33 floats: conflict-free both ways) and then finished with lane = channel:
// residual / c_prev loads and all stores are contiguous runs per pixel instead of one small piece per lane at a
// pixel-sized stride (which costs one LSU wavefront per lane and made the epilogue the slowest stage of the kernel).
constexpr int kStagePitch = 33;
constexpr int kStageFloats = 32 * kStagePitch;             // per epilogue warp
constexpr int kStageBytes = kEpiWarps * kStageFloats * 4;  // 33792

// sigmoid(x) = 1 / (1 + 2^(-x log2 e)) as one MUFU.EX2 and one MUFU.RCP.  __expf wraps the same ex2.approx in a range
// test and two predicated multiplies so that results below 2^-126 come out denormal; after `1 + ...` that makes no
// difference in float32, and those three instructions x 20 transcendentals were ~13 % of the row-wise cell epilogue,
// which runs at latency, not throughput (two epilogue warps per scheduler).
__device__ __forceinline__ float fast_sigmoid(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
__device__ __forceinline__ float fast_tanh(float x) { return fmaf(2.f, fast_sigmoid(2.f * x), -1.f); }

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
__device__ __forceinline__ void store4(void* ptr, size_t plane, int fmt, size_t idx, const float (&v)[4]) {
  if (fmt == RSIS_FMT_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(ptr) + idx) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    // hi = rn(v) two at a time (cvt.rn.bf16x2.f32), lo = rn(v - hi)
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v[0], v[1]), h23 = __floats2bfloat162_rn(v[2], v[3]);
    const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
    const __nv_bfloat162 l01 = __floats2bfloat162_rn(v[0] - __uint_as_float(u01 << 16),
                                                     v[1] - __uint_as_float(u01 & 0xffff0000u));
    const __nv_bfloat162 l23 = __floats2bfloat162_rn(v[2] - __uint_as_float(u23 << 16),
                                                     v[3] - __uint_as_float(u23 & 0xffff0000u));
    __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(ptr);
    *reinterpret_cast<uint2*>(q + idx) = make_uint2(u01, u23);
    *reinterpret_cast<uint2*>(q + idx + plane) =
        make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
  }
}
__device__ __forceinline__ void load4(const View& v, size_t idx, float (&out)[4]) {
  if (v.fmt == RSIS_FMT_F32) {
    const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(v.p) + idx);
    out[0] = t.x; out[1] = t.y; out[2] = t.z; out[3] = t.w;
  } else {
    const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(v.p);
    const uint2 h = *reinterpret_cast<const uint2*>(q + idx);
    const uint2 l = *reinterpret_cast<const uint2*>(q + idx + v.plane);
    out[0] = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
    out[1] = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
    out[2] = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
    out[3] = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
  }
}

// Four adjacent channels of an activation as they sit in memory (no arithmetic, so a prefetch only issues loads):
// float32 -> the four floats; split bf16 -> {hi01, hi23, lo01, lo23}.
__device__ __forceinline__ uint4 load4_raw(const View& v, size_t idx) {
  if (v.fmt == RSIS_FMT_F32) return __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(v.p) + idx));
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(v.p);
  const uint2 h = __ldg(reinterpret_cast<const uint2*>(q + idx));
  const uint2 l = __ldg(reinterpret_cast<const uint2*>(q + idx + v.plane));
  return make_uint4(h.x, h.y, l.x, l.y);
}
__device__ __forceinline__ void decode4_raw(int fmt, const uint4& r, float (&out)[4]) {
  if (fmt == RSIS_FMT_F32) {
    out[0] = __uint_as_float(r.x); out[1] = __uint_as_float(r.y);
    out[2] = __uint_as_float(r.z); out[3] = __uint_as_float(r.w);
  } else {
    out[0] = __uint_as_float(r.x << 16) + __uint_as_float(r.z << 16);
    out[1] = __uint_as_float(r.x & 0xffff0000u) + __uint_as_float(r.z & 0xffff0000u);
    out[2] = __uint_as_float(r.y << 16) + __uint_as_float(r.w << 16);
    out[3] = __uint_as_float(r.y & 0xffff0000u) + __uint_as_float(r.w & 0xffff0000u);
  }
}

// The residual (`out += identity`, torchvision Bottleneck.forward) of one piece, fetched BEFORE the accumulator is
// waited for: loaded inside the finishing loop, every one of its iterations paid a full DRAM latency (the stores of
// the previous iteration may alias, so the compiler cannot hoist the loads) -- 54 us for layer1.conv3 at batch 8
// where the bytes need 12 (profiles/r2a_encoder_ncu_full.md rows 4-5).
template <int PW>
struct ResPiece {
  static constexpr int NIT = 32 / (32 / (PW / 4));
  uint4 v[NIT];
};
template <int PW>
__device__ __forceinline__ void res_prefetch(const UmmaParams& p, ResPiece<PW>& rp, int lane, uint32_t mypix, int col0) {
  constexpr int LPR = PW / 4, RPI = 32 / LPR;
  const int col = col0 + 4 * (lane % LPR);
  const bool col_ok = col < p.Cout;
#pragma unroll
  for (int it = 0; it < 32 / RPI; ++it) {
    const uint32_t pix = __shfl_sync(0xffffffffu, mypix, it * RPI + lane / LPR);
    rp.v[it] = (pix != 0xffffffffu && col_ok) ? load4_raw(p.res, (size_t)pix * p.res_cs + col) : make_uint4(0u, 0u, 0u, 0u);
  }
}

// Finishes one staged piece of a plain convolution: folded BN/bias (+residual) (+ReLU) -> y (and y2).
// stage[row * 33 + c] holds accumulator (row, c); mypix is this lane's row's pixel index (0xffffffff = outside).
// Lane = (row within the iteration, group of 4 adjacent channels): a pixel's PW channels are one contiguous run.
template <int PW>
__device__ __forceinline__ void conv_finish_piece(const UmmaParams& p, const float* stage, int lane, uint32_t mypix,
                                                  int col0, const ResPiece<PW>& rp, const float4 sc, const float4 sh) {
  constexpr int LPR = PW / 4;    // lanes per row, four adjacent columns each
  constexpr int RPI = 32 / LPR;  // rows per iteration
  const int c = 4 * (lane % LPR);
  const int col = col0 + c;
  const bool col_ok = col < p.Cout;  // Cout % 4 == 0
  const float* srow = stage + (lane / LPR) * kStagePitch + c;
#pragma unroll
  for (int it = 0; it < 32 / RPI; ++it) {
    const int row = it * RPI + lane / LPR;
    const uint32_t pix = __shfl_sync(0xffffffffu, mypix, row);
    const float* sp = srow + it * RPI * kStagePitch;
    float v[4];
    v[0] = fmaf(sp[0], sc.x, sh.x);
    v[1] = fmaf(sp[1], sc.y, sh.y);
    v[2] = fmaf(sp[2], sc.z, sh.z);
    v[3] = fmaf(sp[3], sc.w, sh.w);
    if (pix != 0xffffffffu && col_ok) {
      if (p.has_res) {
        float t[4];
        decode4_raw(p.res.fmt, rp.v[it], t);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] += t[j];
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
      }
      store4(p.y, p.y_plane, p.y_fmt, (size_t)pix * p.y_cs + col, v);
      if (p.y2) store4(p.y2, p.y2_plane, p.y2_fmt, (size_t)pix * p.y2_cs + col, v);
    }
  }
}

// ConvLSTM cell pieces (clstm.py:50-58): PW gate columns = PW/4 hidden channels x (in, remember, out, cell).
// Lane = (row within the iteration, hidden channel).  c_prev is fetched BEFORE the accumulator is waited for /
// staged, so its DRAM latency overlaps the MMAs and the TMEM read instead of stalling every iteration.
template <int PW>
struct CellPiece {
  static constexpr int CPP = PW / 4;    // hidden channels in the piece
  static constexpr int RPI = 32 / CPP;  // rows per iteration
  static constexpr int NIT = 32 / RPI;  // iterations
  uint32_t pix[NIT];
  float cp[NIT];
  float4 pre[NIT];
};

template <int PW>
__device__ __forceinline__ void cell_prefetch(const UmmaParams& p, CellPiece<PW>& cpz, int lane, uint32_t mypix,
                                              int col0) {
  using CP = CellPiece<PW>;
  const int Ch = p.Cout >> 2;
  const int chg = (col0 >> 2) + lane % CP::CPP;
  const bool ch_ok = chg < Ch;
#pragma unroll
  for (int it = 0; it < CP::NIT; ++it) {
    const int row = it * CP::RPI + lane / CP::CPP;
    const uint32_t pix = __shfl_sync(0xffffffffu, mypix, row);
    cpz.pix[it] = ch_ok ? pix : 0xffffffffu;
    cpz.cp[it] = (p.c_prev && cpz.pix[it] != 0xffffffffu) ? __ldg(p.c_prev + (size_t)pix * Ch + chg) : 0.f;
    cpz.pre[it] = (p.preact && cpz.pix[it] != 0xffffffffu)
                      ? __ldg(reinterpret_cast<const float4*>(p.preact + ((size_t)pix * Ch + chg) * 4))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// The side feature (global nn.MaxPool2d of model.py:143) is a running max per lane, merged across the lanes of a
// channel at the end.
template <int PW>
__device__ __forceinline__ void cell_finish_piece(const UmmaParams& p, const float* stage, int lane,
                                                  const CellPiece<PW>& cpz, int myimg, int col0, int seg) {
  using CP = CellPiece<PW>;
  const int Ch = p.Cout >> 2;
  const int ch = lane % CP::CPP;
  const int chg = (col0 >> 2) + ch;
  const bool ch_ok = chg < Ch;
  float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
  if (ch_ok) {
    sc = __ldg(reinterpret_cast<const float4*>(p.scale + col0 + 4 * ch));
    sh = __ldg(reinterpret_cast<const float4*>(p.shift + col0 + 4 * ch));
  }
  uint32_t best = 0u;  // key 0 sorts below every float
#pragma unroll
  for (int it = 0; it < CP::NIT; ++it) {
    const int row = it * CP::RPI + lane / CP::CPP;
    const uint32_t pix = cpz.pix[it];
    const float* g = stage + row * kStagePitch + 4 * ch;
    const float gi = fast_sigmoid(fmaf(g[0], sc.x, sh.x) + cpz.pre[it].x);
    const float gf = fast_sigmoid(fmaf(g[1], sc.y, sh.y) + cpz.pre[it].y);
    const float go = fast_sigmoid(fmaf(g[2], sc.z, sh.z) + cpz.pre[it].z);
    const float gg = fast_tanh(fmaf(g[3], sc.w, sh.w) + cpz.pre[it].w);
    const float cv = fmaf(gf, cpz.cp[it], gi * gg);
    const float hv = go * fast_tanh(cv);
    if (pix == 0xffffffffu) continue;
    const size_t idx = (size_t)pix * Ch + chg;
    p.c_out[idx] = cv;
    p.h_out[idx] = hv;
    if (p.h_split) {
      __nv_bfloat16 hi, lo;
      split_bf16(hv, hi, lo);
      const size_t k = (size_t)pix * p.hs_cs + chg;
      p.h_split[k] = hi;
      p.h_split[k + p.hs_plane] = lo;
    }
    if (p.side_max) {
      const uint32_t key = float_to_key(hv);
      if (seg == 32) {
        best = key > best ? key : best;
      } else {  // maps smaller than a warp's 32 rows: rows of several images share the warp
        const int img = (int)(pix / (uint32_t)(p.Ho * p.Wo));
        atomicMax(p.side_max + (size_t)img * p.side_stride + p.side_offset + chg, key);
      }
    }
  }
  if (p.side_max && seg == 32) {
#pragma unroll
    for (int s = CP::CPP; s < 32; s <<= 1) {
      const uint32_t o = __shfl_xor_sync(0xffffffffu, best, s);
      best = o > best ? o : best;
    }
    const int img0 = __shfl_sync(0xffffffffu, myimg, 0);
    if (lane < CP::CPP && ch_ok && best != 0u)
      atomicMax(p.side_max + (size_t)img0 * p.side_stride + p.side_offset + chg, best);
  }
}

// Row (= TMEM lane = output pixel of the tile) -> pixel geometry.
struct RowGeom {
  bool ok;
  size_t pix;
  int img;
};
__device__ __forceinline__ RowGeom row_geom(const UmmaParams& p, int row, int tw, int th, int ti) {
  const int wl = row % p.BW;
  const int hl = (row / p.BW) % p.BH;
  const int il = row / (p.BW * p.BH);
  const int wo = tw * p.BW + wl, ho = th * p.BH + hl, img = ti * p.BI + il;
  RowGeom g;
  g.ok = wo < p.Wo && ho < p.Ho && img < p.N;
  g.pix = ((size_t)img * p.Ho + ho) * p.Wo + wo;
  g.img = img;
  return g;
}

// (tile, K slice) of a work unit and (n tile, w, h, image tile) of a tile, with the launch's fast dividers.
struct TileCoord {
  int nt, tw, th, ti;
};
__device__ __forceinline__ void decode_work(const UmmaParams& p, int work, int& tile, int& ks) {
  tile = (int)fdiv((uint32_t)work, p.dks);
  ks = work - tile * p.ksplit;
}
__device__ __forceinline__ TileCoord decode_tile(const UmmaParams& p, int tile) {
  TileCoord t;
  uint32_t mt = fdiv((uint32_t)tile, p.dn);
  t.nt = tile - (int)mt * p.tiles_n;
  uint32_t q = fdiv(mt, p.dw);
  t.tw = (int)mt - (int)q * p.tiles_w;
  mt = q;
  q = fdiv(mt, p.dh);
  t.th = (int)mt - (int)q * p.tiles_h;
  t.ti = (int)q;
  return t;
}
// A thread's position inside the pixel box never changes: computed once per thread (three real divisions).
struct RowPos {
  int wl, hl, il;
};
__device__ __forceinline__ RowPos row_pos(const UmmaParams& p, int row) {
  RowPos r;
  r.wl = row % p.BW;
  r.hl = (row / p.BW) % p.BH;
  r.il = row / (p.BW * p.BH);
  return r;
}
__device__ __forceinline__ RowGeom row_geom(const UmmaParams& p, const RowPos& r, const TileCoord& t) {
  const int wo = t.tw * p.BW + r.wl, ho = t.th * p.BH + r.hl, img = t.ti * p.BI + r.il;
  RowGeom g;
  g.ok = wo < p.Wo && ho < p.Ho && img < p.N;
  g.pix = ((size_t)img * p.Ho + ho) * p.Wo + wo;
  g.img = img;
  return g;
}

#ifdef RSIS_DEBUG_TIMING
#ifndef RSIS_TRACE_BLOCK
#define RSIS_TRACE_BLOCK 0  // the block whose stamps go to the per-launch trace (-DRSIS_TRACE_BLOCK=100: a block that
#endif                      // cannot be resident before the previous kernel's CTAs leave)
__device__ __forceinline__ void stamp(const UmmaParams& p, int slot) {
  if (blockIdx.x == 0 && p.counters) {
    // SM cycle counter (all stamps of the table come from block 0, i.e. one SM): ~20 cycles, where %globaltimer costs
    // hundreds of nanoseconds and visibly stretched the roles it was stamping; readers divide by the SM clock
    const unsigned long long t = (unsigned long long)clock64();
    reinterpret_cast<unsigned long long*>(p.counters)[256 + slot] = t;
  }
  // per-launch trace (rsis_debug_trace): block 0's first sixteen stamps go to the launch's own row, so that launches
  // overlapping under programmatic dependent launch do not overwrite each other
  if (blockIdx.x == RSIS_TRACE_BLOCK && p.trace && slot < 16) p.trace[2 + slot] = (unsigned long long)clock64();
}
__device__ __forceinline__ void trace_begin(const UmmaParams& p) {
  if (threadIdx.x == 0 && p.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(p.trace + 0, ~t);  // earliest CTA start, inverted
    if (blockIdx.x == RSIS_TRACE_BLOCK) {
      p.trace[18] = t;
      p.trace[19] = (unsigned long long)clock64();
    }
  }
}
__device__ __forceinline__ void trace_end(const UmmaParams& p) {
  if (threadIdx.x == 0 && p.trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(p.trace + 1, t);  // latest CTA end
    atomicMax(p.trace + 22, ~t);  // earliest CTA end, inverted
    atomicAdd(p.trace + 23, 1ull);  // CTAs
    if (blockIdx.x == RSIS_TRACE_BLOCK) {
      p.trace[20] = t;
      p.trace[21] = (unsigned long long)clock64();
    }
  }
}
#define STAMP(slot) stamp(p, slot)
// per-tile timeline of block 0: role 0 = A issued, 1 = MMA sees A, 2 = MMAs issued, 3 = epilogue sees accumulator,
// 4 = epilogue released the accumulator; up to 16 tiles each
#define STAMP_T(role, iter) do { if ((iter) < 16) stamp(p, 16 + (role) * 16 + (iter)); } while (0)
#else
#define STAMP(slot)
#define STAMP_T(role, iter)
#endif

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// Epilogue warps: accumulator pieces (32 rows x PW columns) -> staged transpose -> finish.  With split-K the CTA first
// parks its partial accumulator in the L2-resident scratch (TMEM-native order: for a fixed column the 32 lanes of a
// warp write 32 consecutive floats), waits for the other K slices of its tile, then reduces and finishes its share.
template <bool CELL, int PW, bool SPLIT, bool CS = false>  // CS: the cluster split-K variant (its own instantiations: the
                                                           // extra epilogue costs the ordinary kernels registers)
__device__ __forceinline__ void epilogue_role(const UmmaParams& p, uint32_t tmem_base, uint32_t tfull0,
                                              uint32_t tempty0, float* stage, int warp, int lane, const int bid,
                                              const int nblk, uint32_t smem_x = 0, uint32_t xfull = 0) {
  const int quarter = warp & 3;       // TMEM lane quarter this warp may read
  const int half = warp >> 2;         // which pieces (even / odd) of the quarter this warp handles
  const int rows_per_img = p.BW * p.BH;
  const int seg = rows_per_img < 32 ? rows_per_img : 32;
  const int npc = p.BN / PW;          // output pieces per row quarter
  const int ncols = p.stacked ? 2 * p.BN : p.BN;
  const int num_work = p.num_tiles * p.ksplit;
  const RowPos rpos = row_pos(p, quarter * 32 + lane);
  int acc = 0;
  uint32_t acc_phase = 0;
  CellPiece<PW> cpz_cur;
  ResPiece<PW> res_cur;
  for (int work = bid; work < num_work; work += nblk) {
    int tile, ks;
    decode_work(p, work, tile, ks);
    const TileCoord tc = decode_tile(p, tile);
    const int nt = tc.nt, tw = tc.tw, th = tc.th, ti = tc.ti;
    const RowGeom g = row_geom(p, rpos, tc);
    const uint32_t mypix = g.ok ? (uint32_t)g.pix : 0xffffffffu;
    if constexpr (CELL && !SPLIT) {
      // the very first piece of this warp; afterwards every piece's c_prev / hoisted-gate loads are issued one piece
      // ahead (rolling, across tiles), so their DRAM latency hides behind the previous piece instead of stalling it
      if (work == bid && half < npc) cell_prefetch<PW>(p, cpz_cur, lane, mypix, nt * p.BN + PW * half);
    }
    if constexpr (!CELL && !SPLIT) {
      if (p.has_res && work == bid && half < npc) res_prefetch<PW>(p, res_cur, lane, mypix, nt * p.BN + PW * half);
    }
    mbar_wait(tfull0 + 8 * acc, acc_phase);
    tc_fence_after();
    if (threadIdx.x == 0) STAMP(7);
    if (threadIdx.x == 0) STAMP_T(3, (work - bid) / nblk);
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kStageCols;
    // Cluster split-K (csplit; a launch is ONE wave: this CTA has exactly one work unit = one K half of a tile).  The two
    // CTAs of the cluster split the tile's OUTPUT columns too: rank r finishes pieces [r * npc / 2, (r + 1) * npc / 2).
    // Each thread (a) tells the partner that this CTA's MMAs are done -- the exchange buffer lies on the operand stages --
    // (b) once the partner has said the same, writes its partial sums of the PARTNER's pieces, already summed over the
    // stacked halves, row by row (thread = row, 16-byte stores) into the partner's buffer through distributed shared
    // memory, (c) release-arrives on the partner's barrier, (d) waits for the partner's 256 arrivals and finishes its own
    // pieces with the partner's partial sums added.  No scratch in L2, no counters, and the epilogue work is halved.
    if constexpr (CS && !CELL && !SPLIT) {
      {
        const uint32_t rank = cluster_ctarank(), peer = rank ^ 1u;
        const uint32_t xrow = smem_x + (uint32_t)(quarter * 32 + lane) * (uint32_t)p.x_pitch * 4u;
        const uint32_t xrow_remote = mapa_shared(xrow, peer);
        const int own0 = (int)rank * (npc >> 1), own1 = own0 + (npc >> 1);  // npc is even (BN >= 64)
        mbar_arrive_remote(mapa_shared(xfull + 8, peer));  // (a) my operand stages are dead
        mbar_wait_cluster(xfull + 8, 0);                    // (b) ... and so are the partner's
        for (int j = half; j < npc; j += 2) {
          if (j >= own0 && j < own1) continue;
          uint32_t r[PW];
          tmem_ld_piece<PW>(taddr + PW * j, r);
          if (p.stacked) {
            uint32_t r2[PW];
            tmem_ld_piece<PW>(taddr + p.BN + PW * j, r2);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < PW; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
          } else {
            tmem_ld_wait();
          }
#pragma unroll
          for (int e = 0; e < PW; e += 4)
            st_cluster_v4(xrow_remote + (uint32_t)(PW * j + e) * 4u, __uint_as_float(r[e]), __uint_as_float(r[e + 1]),
                          __uint_as_float(r[e + 2]), __uint_as_float(r[e + 3]));
        }
        mbar_arrive_remote(mapa_shared(xfull, peer));       // (c)
        bool have_partner = false;
        for (int j = half; j < npc; j += 2) {
          if (j < own0 || j >= own1) continue;
          const int col0 = nt * p.BN + PW * j;
          if (col0 >= p.Cout) break;
          ResPiece<PW> rp;
          if (p.has_res) res_prefetch<PW>(p, rp, lane, mypix, col0);
          float4 sc4 = make_float4(0.f, 0.f, 0.f, 0.f), sh4 = sc4;
          const int colq = col0 + 4 * (lane % (PW / 4));
          if (colq < p.Cout) {
            sc4 = __ldg(reinterpret_cast<const float4*>(p.scale + colq));
            sh4 = __ldg(reinterpret_cast<const float4*>(p.shift + colq));
          }
          uint32_t r[PW];
          tmem_ld_piece<PW>(taddr + PW * j, r);
          if (p.stacked) {
            uint32_t r2[PW];
            tmem_ld_piece<PW>(taddr + p.BN + PW * j, r2);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < PW; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
          } else {
            tmem_ld_wait();
          }
          if (!have_partner) {
            mbar_wait_cluster(xfull, 0);                    // (d)
            have_partner = true;
          }
#pragma unroll
          for (int e = 0; e < PW; e += 4) {
            float4 q;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w)
                         : "r"(xrow + (uint32_t)(PW * j + e) * 4u));
            r[e] = __float_as_uint(__uint_as_float(r[e]) + q.x);
            r[e + 1] = __float_as_uint(__uint_as_float(r[e + 1]) + q.y);
            r[e + 2] = __float_as_uint(__uint_as_float(r[e + 2]) + q.z);
            r[e + 3] = __float_as_uint(__uint_as_float(r[e + 3]) + q.w);
          }
#pragma unroll
          for (int e = 0; e < PW; ++e) stage[lane * kStagePitch + e] = __uint_as_float(r[e]);
          __syncwarp();
          conv_finish_piece<PW>(p, stage, lane, mypix, col0, rp, sc4, sh4);
          __syncwarp();
        }
        tc_fence_before();
        mbar_arrive(tempty0 + 8 * acc);
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
        continue;
      }
    }
    if constexpr (!SPLIT) {
      for (int j = half; j < npc; j += 2) {
        const int col0 = nt * p.BN + PW * j;
        if (col0 >= p.Cout) break;
        CellPiece<PW> cpz_next;
        if constexpr (CELL) {
          const int coln = col0 + 2 * PW;
          if (j + 2 < npc && coln < p.Cout) {
            cell_prefetch<PW>(p, cpz_next, lane, mypix, coln);
          } else if (work + nblk < num_work) {
            int tile2, ks2;
            decode_work(p, work + nblk, tile2, ks2);
            const TileCoord tc2 = decode_tile(p, tile2);
            const int nt2 = tc2.nt;
            const RowGeom g2 = row_geom(p, rpos, tc2);
            cell_prefetch<PW>(p, cpz_next, lane, g2.ok ? (uint32_t)g2.pix : 0xffffffffu, nt2 * p.BN + PW * half);
          }
        }
        ResPiece<PW> res_next;
        if constexpr (!CELL) {
          if (p.has_res) {  // the next piece's residual (of this tile, or the first piece of this CTA's next tile)
            const int coln = col0 + 2 * PW;
            if (j + 2 < npc && coln < p.Cout) {
              res_prefetch<PW>(p, res_next, lane, mypix, coln);
            } else if (work + nblk < num_work) {
              int tile2, ks2;
              decode_work(p, work + nblk, tile2, ks2);
              const TileCoord tc2 = decode_tile(p, tile2);
              const RowGeom g2 = row_geom(p, rpos, tc2);
              res_prefetch<PW>(p, res_next, lane, g2.ok ? (uint32_t)g2.pix : 0xffffffffu, tc2.nt * p.BN + PW * half);
            }
          }
        }
        // the piece's folded affine first (its L1 / L2 latency hides behind the tensor-memory loads instead of following
        // the transpose), then BOTH tensor-memory loads of a stacked accumulator before one wait
        float4 sc4 = make_float4(0.f, 0.f, 0.f, 0.f), sh4 = sc4;
        if constexpr (!CELL) {
          const int colq = col0 + 4 * (lane % (PW / 4));
          if (colq < p.Cout) {
            sc4 = __ldg(reinterpret_cast<const float4*>(p.scale + colq));
            sh4 = __ldg(reinterpret_cast<const float4*>(p.shift + colq));
          }
        }
        uint32_t r[PW];
        tmem_ld_piece<PW>(taddr + PW * j, r);
        if (p.stacked) {
          uint32_t r2[PW];
          tmem_ld_piece<PW>(taddr + p.BN + PW * j, r2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < PW; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
        } else {
          tmem_ld_wait();
        }
#pragma unroll
        for (int e = 0; e < PW; ++e) stage[lane * kStagePitch + e] = __uint_as_float(r[e]);
        __syncwarp();
        if constexpr (CELL) {
          cell_finish_piece<PW>(p, stage, lane, cpz_cur, g.img, col0, seg);
          cpz_cur = cpz_next;
        } else {
          conv_finish_piece<PW>(p, stage, lane, mypix, col0, res_cur, sc4, sh4);
          if (p.has_res) res_cur = res_next;
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(tempty0 + 8 * acc);
      if (threadIdx.x == 0) STAMP_T(4, (work - bid) / nblk);
    } else {
      // ---- split-K (1): park this CTA's partial accumulator in the scratch as [row][column] (staged transpose,
      // so each row's 32 columns are one 128-byte store)
      {
        float* dst = p.scratch + ((size_t)work * kBM + quarter * 32) * ncols;
        for (int j = half; j < (ncols >> 5); j += 2) {
          uint32_t r[32];
          tmem_ld32(taddr + 32 * j, r);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) stage[lane * kStagePitch + e] = __uint_as_float(r[e]);
          __syncwarp();
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) __stcg(dst + (size_t)rr * ncols + 32 * j + lane, stage[rr * kStagePitch + lane]);
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(tempty0 + 8 * acc);
      if (threadIdx.x == 0) STAMP(8);
      // ---- (2) wait until all K slices of the tile are parked
      __threadfence();
      epi_bar();
      if (threadIdx.x == 0) {
        atomicAdd(p.counters + 2 * tile, 1u);
        unsigned seen;
        const long long t0 = clock64();
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.counters + 2 * tile) : "memory");
          if (clock64() - t0 > 4000000000LL) {
            printf("rsis_b200 conv_umma: split-K wait timed out (block %d tile %d seen %u of %d)\n", bid, tile,
                   seen, p.ksplit);
            __trap();
          }
        } while (seen < (unsigned)p.ksplit);
        STAMP(9);
      }
      epi_bar();
      // ---- (3) reduce + finish.  Unit = 4 rows x 32 columns: lane = (row, 4 adjacent columns), every load and
      // store is a 128-byte (or per-pixel contiguous) run.  Units are dealt round-robin to the 8 * ksplit warps.
      {
        const int ncg = p.BN >> 5;
        const int units = (kBM / 4) * ncg;
        const int Ch = p.Cout >> 2;
        struct Unit {
          RowGeom g;
          int col;
          bool ok;
          float cprev;
          float4 pre;
          float4 v;
          uint4 res;
        };
        // loads of one unit: issued for two units before either is finished, so their L2 / DRAM latencies overlap
        auto unit_load = [&](int u) {
          Unit t;
          const int cg = u % ncg;
          const int row = 4 * (u / ncg) + (lane >> 3);
          const int c4 = lane & 7;
          t.col = nt * p.BN + 32 * cg + 4 * c4;
          t.g = row_geom(p, row, tw, th, ti);
          t.ok = t.g.ok && t.col < p.Cout;
          t.cprev = 0.f;
          if constexpr (CELL) {
            if (t.ok && p.c_prev) t.cprev = __ldg(p.c_prev + t.g.pix * Ch + (t.col >> 2));
            t.pre = (t.ok && p.preact) ? __ldg(reinterpret_cast<const float4*>(p.preact + t.g.pix * p.Cout + t.col))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          t.res = make_uint4(0u, 0u, 0u, 0u);
          if constexpr (!CELL) {
            if (t.ok && p.has_res) t.res = load4_raw(p.res, t.g.pix * p.res_cs + t.col);
          }
          t.v = make_float4(0.f, 0.f, 0.f, 0.f);
          const float* src = p.scratch + ((size_t)(tile * p.ksplit) * kBM + row) * ncols + 32 * cg + 4 * c4;
          const size_t sstride = (size_t)kBM * ncols;
#pragma unroll 8
          for (int s2 = 0; s2 < p.ksplit; ++s2) {
            const float4 a4 = __ldcg(reinterpret_cast<const float4*>(src + s2 * sstride));
            t.v.x += a4.x; t.v.y += a4.y; t.v.z += a4.z; t.v.w += a4.w;
            if (p.stacked) {
              const float4 b4 = __ldcg(reinterpret_cast<const float4*>(src + s2 * sstride + p.BN));
              t.v.x += b4.x; t.v.y += b4.y; t.v.z += b4.z; t.v.w += b4.w;
            }
          }
          return t;
        };
        auto unit_finish = [&](const Unit& t) {
          const int col = t.col;
          const float4 v = t.v;
          float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sh = sc;
          if (col < p.Cout) {
            sc = __ldg(reinterpret_cast<const float4*>(p.scale + col));
            sh = __ldg(reinterpret_cast<const float4*>(p.shift + col));
          }
          if constexpr (CELL) {
            const int chg = col >> 2;
            const float gi = fast_sigmoid(fmaf(v.x, sc.x, sh.x) + t.pre.x);
            const float gf = fast_sigmoid(fmaf(v.y, sc.y, sh.y) + t.pre.y);
            const float go = fast_sigmoid(fmaf(v.z, sc.z, sh.z) + t.pre.z);
            const float gg = fast_tanh(fmaf(v.w, sc.w, sh.w) + t.pre.w);
            const float cv = fmaf(gf, t.cprev, gi * gg);
            const float hv = go * fast_tanh(cv);
            uint32_t key = 0u;
            if (t.ok) {
              const size_t idx = t.g.pix * Ch + chg;
              p.c_out[idx] = cv;
              p.h_out[idx] = hv;
              if (p.h_split) {
                __nv_bfloat16 hi, lo;
                split_bf16(hv, hi, lo);
                const size_t k = t.g.pix * p.hs_cs + chg;
                p.h_split[k] = hi;
                p.h_split[k + p.hs_plane] = lo;
              }
              key = float_to_key(hv);
            }
            if (p.side_max) {
              if ((rows_per_img & 3) == 0) {  // the unit's 4 rows belong to one image: merge them first
                uint32_t o = __shfl_xor_sync(0xffffffffu, key, 8);
                key = o > key ? o : key;
                o = __shfl_xor_sync(0xffffffffu, key, 16);
                key = o > key ? o : key;
                const int img0 = __shfl_sync(0xffffffffu, t.g.img, 0);
                if (lane < 8 && key != 0u)
                  atomicMax(p.side_max + (size_t)img0 * p.side_stride + p.side_offset + chg, key);
              } else if (key != 0u) {
                atomicMax(p.side_max + (size_t)t.g.img * p.side_stride + p.side_offset + chg, key);
              }
            }
          } else {
            float o[4];
            o[0] = fmaf(v.x, sc.x, sh.x);
            o[1] = fmaf(v.y, sc.y, sh.y);
            o[2] = fmaf(v.z, sc.z, sh.z);
            o[3] = fmaf(v.w, sc.w, sh.w);
            if (t.ok) {
              if (p.has_res) {
                float r4[4];
                decode4_raw(p.res.fmt, t.res, r4);
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] += r4[e];
              }
              if (p.relu) {
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
              }
              store4(p.y, p.y_plane, p.y_fmt, t.g.pix * p.y_cs + col, o);
              if (p.y2) store4(p.y2, p.y2_plane, p.y2_fmt, t.g.pix * p.y2_cs + col, o);
            }
          }
        };
        const int ustride = kEpiWarps * p.ksplit;
        for (int u = ks * kEpiWarps + warp; u < units; u += 2 * ustride) {
          const Unit t0 = unit_load(u);
          if (u + ustride < units) {
            const Unit t1 = unit_load(u + ustride);
            unit_finish(t0);
            unit_finish(t1);
          } else {
            unit_finish(t0);
          }
        }
      }
      epi_bar();
      if (threadIdx.x == 0) STAMP(10);
      if (threadIdx.x == 0) {
        // the last CTA of the tile to finish re-arms the counters for the next launch that uses this scratch
        if (atomicAdd(p.counters + 2 * tile + 1, 1u) == (unsigned)p.ksplit - 1u) {
          p.counters[2 * tile] = 0u;
          p.counters[2 * tile + 1] = 0u;
        }
      }
    }
    if (++acc == kAccStages) {
      acc = 0;
      acc_phase ^= 1u;
    }
  }
}

// Bounded wait without the diagnostic printf (used inside fully unrolled issue sequences, where code size matters).
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// All MMAs of one HALO activation item whose nine weight taps sit in slots 0..8 of the B ring (resident weights): a
// straight line of up to 9 x KS x {2, 3} UTCHMMAs whose descriptors are `uniform base + compile-time offset`, so they
// issue back to back from uniform registers (the generic loop spends ~50 dependent instructions per tap in one thread,
// which at 4 MMAs of 48 cycles per tap is slower than the tensor pipe: profiles/r2i_group_stamps.txt).
// tap (kh, kw) = the activation tile shifted by kh*10 + kw rows of 128 bytes = (kh*10 + kw) * 8 descriptor units.
template <int KS, int PASSES, int ROW16 = 8>  // PASSES: 1 = single-pass bf16, 2 = stacked weight planes, 3 = three products;
                                              // ROW16: bytes / 16 of an activation row (8: SWIZZLE_128B, 4: SWIZZLE_64B)
__device__ __forceinline__ void issue_halo_resident(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t bdesc0,
                                                    uint32_t b_stage16, uint32_t b_plane16, uint32_t idesc,
                                                    uint32_t idesc_lo, uint32_t accumulate, bool wait_b, uint32_t bfull0,
                                                    int bg) {
  // b_stage16: one tap's weights (a K block); the taps arrive in boxes of `bg`, one barrier per box
#pragma unroll
  for (int bi = 0; bi < 9; ++bi) {
    if (wait_b && (bg == 1 || bi % bg == 0)) {
      mbar_wait_lean(bfull0 + 8 * (bg == 1 ? bi : bi / bg), 0);
      tc_fence_after();
    }
    const uint64_t toff = (uint64_t)(((bi / 3) * (kHaloBW + 2) + (bi % 3)) * ROW16);
    const uint64_t b = bdesc0 + (uint64_t)((uint32_t)bi * b_stage16);
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const uint64_t adv = (uint64_t)(2 * k);  // 16 bf16 = 32 bytes along K
      const uint32_t first = (bi == 0 && k == 0) ? accumulate : 1u;
      if (PASSES == 1) {
        umma_bf16(d, a_hi + toff + adv, b + adv, idesc, first);
      } else if (PASSES == 2) {
        umma_bf16(d, a_hi + toff + adv, b + adv, idesc, first);
        umma_bf16(d, a_lo + toff + adv, b + adv, idesc_lo, 1u);
      } else {
        umma_bf16(d, a_hi + toff + adv, b + adv, idesc, first);
        umma_bf16(d, a_hi + toff + adv, b + b_plane16 + adv, idesc, 1u);
        umma_bf16(d, a_lo + toff + adv, b + adv, idesc, 1u);
      }
    }
  }
}

// The same for a 33..48-channel source staged as a 32-channel SWIZZLE_64B part (K steps 0, 1) and a 16-channel
// SWIZZLE_32B part (K step 2): the activation descriptors of the two parts differ in layout type, row size (tap shift
// of 4 / 2 descriptor units per row) and group stride; the weight tile is the ordinary 64-channel SWIZZLE_128B one.
__device__ __forceinline__ void issue_halo_resident_split48(uint32_t d, uint64_t a64_hi, uint64_t a64_lo, uint64_t a32_hi,
                                                            uint64_t a32_lo, uint64_t bdesc0, uint32_t b_stage16,
                                                            uint32_t idesc, uint32_t idesc_lo, uint32_t accumulate, bool wait_b,
                                                            uint32_t bfull0, int bg) {
#pragma unroll
  for (int bi = 0; bi < 9; ++bi) {
    if (wait_b && (bg == 1 || bi % bg == 0)) {
      mbar_wait_lean(bfull0 + 8 * (bg == 1 ? bi : bi / bg), 0);
      tc_fence_after();
    }
    const int row = (bi / 3) * (kHaloBW + 2) + (bi % 3);
    const uint64_t b = bdesc0 + (uint64_t)((uint32_t)bi * b_stage16);
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const uint64_t adv = (uint64_t)(2 * k);
      umma_bf16(d, a64_hi + (uint64_t)(row * 4) + adv, b + adv, idesc, (bi == 0 && k == 0) ? accumulate : 1u);
      umma_bf16(d, a64_lo + (uint64_t)(row * 4) + adv, b + adv, idesc_lo, 1u);
    }
    umma_bf16(d, a32_hi + (uint64_t)(row * 2), b + 4, idesc, 1u);
    umma_bf16(d, a32_lo + (uint64_t)(row * 2), b + 4, idesc_lo, 1u);
  }
}

// The MMAs of one (tap, chunk) K block: KS steps of 16 channels, compile-time descriptor offsets.
template <int KS, int PASSES>  // PASSES: 1 = single-pass bf16, 2 = stacked weight planes, 3 = three products
__device__ __forceinline__ void issue_kblock(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b, uint32_t b_plane16,
                                             uint32_t idesc, uint32_t idesc_lo, uint32_t accumulate) {
#pragma unroll
  for (int k = 0; k < KS; ++k) {
    const uint64_t adv = (uint64_t)(2 * k);  // 16 bf16 = 32 bytes along K inside the swizzle row
    const uint32_t first = k == 0 ? accumulate : 1u;
    if (PASSES == 1) {
      umma_bf16(d, a_hi + adv, b + adv, idesc, first);
    } else if (PASSES == 2) {
      umma_bf16(d, a_hi + adv, b + adv, idesc, first);
      umma_bf16(d, a_lo + adv, b + adv, idesc_lo, 1u);
    } else {
      umma_bf16(d, a_hi + adv, b + adv, idesc, first);
      umma_bf16(d, a_hi + adv, b + b_plane16 + adv, idesc, 1u);
      umma_bf16(d, a_lo + adv, b + adv, idesc, 1u);
    }
  }
}

// All MMAs of one HALO activation item (a 64-channel chunk, four K steps per tap) whose weights STREAM through the B ring
// in boxes of three taps: per box one barrier wait, 3 x 4 x PASSES straight-line UTCHMMAs and one commit.  The issuing
// thread spends ~550 cycles of barrier / descriptor / constant-load overhead per wait-issue-commit round (ncu source
// view + in-kernel stamps, profiles/r2z_*): with one round per tap (8 MMAs of 48 cycles) the tensor pipe idled 60 % of
// the main loop of layer3.conv2; a round per 24 MMAs keeps it fed.
template <int PASSES>  // 2 = stacked weight planes, 3 = three products
__device__ __forceinline__ void issue_halo_stream3(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t bdesc0,
                                                   uint32_t b_stage16, uint32_t b_chunk16, uint32_t b_plane16,
                                                   uint32_t idesc, uint32_t idesc_lo, uint32_t accumulate, uint32_t bfull0, uint32_t bempty0,
                                                   int& bs, uint32_t& bph, int b_stages) {
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    mbar_wait_lean(bfull0 + 8 * bs, bph);
    tc_fence_after();
    const uint64_t b0 = bdesc0 + (uint64_t)((uint32_t)bs * b_stage16);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const uint64_t toff = (uint64_t)((g * (kHaloBW + 2) + j) * 8);
      const uint64_t b = b0 + (uint64_t)((uint32_t)j * b_chunk16);
#pragma unroll
      for (int k = 0; k < kBK / 16; ++k) {
        const uint64_t adv = (uint64_t)(2 * k);
        const uint32_t first = (g == 0 && j == 0 && k == 0) ? accumulate : 1u;
        if (PASSES == 2) {
          umma_bf16(d, a_hi + toff + adv, b + adv, idesc, first);
          umma_bf16(d, a_lo + toff + adv, b + adv, idesc_lo, 1u);
        } else {
          umma_bf16(d, a_hi + toff + adv, b + adv, idesc, first);
          umma_bf16(d, a_hi + toff + adv, b + b_plane16 + adv, idesc, 1u);
          umma_bf16(d, a_lo + toff + adv, b + adv, idesc, 1u);
        }
      }
    }
    umma_commit(bempty0 + 8 * bs);  // weight slot free once these MMAs have read it
    if (++bs == b_stages) {
      bs = 0;
      bph ^= 1u;
    }
  }
}

// ---- ConvLSTM epilogue, row-wise (clstm.py:50-58) ------------------------------------------------------------------
// The accumulator leaves TMEM with thread = pixel (TMEM lane) and registers = gate columns in (hidden channel, gate)
// order, so ONE thread holds i, f, o, g of a hidden channel of its pixel: the gate math needs no data exchange at all.
// A unit = 16 gate columns = 4 hidden channels of 32 pixels.  Per thread and unit: one 16-byte load of c_prev, four of
// the hoisted gate share (64 contiguous bytes), the stores of c, h (16 bytes each) and of h as split bf16 (2 x 8 bytes),
// all independent across the 4 channels -- instruction-level parallelism instead of the staged transpose's dependent
// STS -> sync -> LDS chain (with 2-3 warps per scheduler that chain, not the instruction count, set the epilogue time:
// in-kernel stamps, profiles/r2f_group_stamps.txt).  The two warps of a TMEM lane quarter split a tile's units; the loads
// of the next unit (possibly of the next tile) are issued before the current one is finished, and the accumulator stage
// is handed back to the MMA warp as soon as its last unit is in registers.
struct CellUnitIn {
  float4 pre[4];
  float4 cp;
};

__device__ __forceinline__ void cell_unit_load(const UmmaParams& p, CellUnitIn& in, uint32_t pix, int chg0) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const int Ch = p.Cout >> 2;
  const bool ok = pix != 0xffffffffu && chg0 < Ch;
  const size_t base = (size_t)pix * Ch + chg0;
  in.cp = (ok && p.c_prev) ? __ldg(reinterpret_cast<const float4*>(p.c_prev + base)) : z;
  if (p.pre_tma) return;  // the gate share comes through shared memory (cell_rows_epilogue)
#pragma unroll
  for (int j = 0; j < 4; ++j)
    in.pre[j] = (ok && p.preact) ? __ldg(reinterpret_cast<const float4*>(p.preact + (base + j) * 4)) : z;
}

// pre_tma: a thread = pixel load of the hoisted gate share touches one 128-byte line per lane (32 L1 wavefronts per
// instruction), and L1 shares its data path with the shared-memory operand reads of the tensor core -- measured: the
// MMAs of a 32-gate-column level ran at 74 ns instead of 26 (profiles/r2k_group_stamps.txt).  With pre_tma the A
// producer warp brings the tile's [128 pixels][32 columns] fp32 boxes in by TMA (SWIZZLE_128B, two stages), and a thread
// reads its own 128-byte row with conflict-free 16-byte shared-memory loads (chunk ^ (row & 7)).
__device__ __forceinline__ void cell_rows_epilogue(const UmmaParams& p, uint32_t tmem_base, uint32_t tfull0,
                                                   uint32_t tempty0, int warp, int lane, const int bid, const int nblk,
                                                   const int wpq) {
  // `wpq` epilogue warps per TMEM lane quarter (2 in the 8-warp kernels, 3 in the grouped launch).  The flat stream of
  // units f = tile_seq * nu + u of a quarter is dealt round-robin to them: a unit runs at LATENCY here (about 450
  // dependent-ish instructions: address arithmetic, two tensor-memory loads, 20 transcendentals, stores, the max-pool
  // reduction -- ~2 us with two such warps per scheduler, ncu source view: profiles/r2bu_*), so a third warp per quarter
  // is worth more than any trimming of the unit.
  const int quarter = warp & 3, slot = warp >> 2;
  const int Ch = p.Cout >> 2;
  const int nu = p.BN >> 4;                  // units per output-channel tile (a power of two: BN = 32 .. 256)
  const int nu_shift = 31 - __clz(nu);
  const bool warp_one_image = p.BW * p.BH >= 32;  // the 32 rows of a warp lie in one image
  const int num_work = p.num_tiles;               // no split-K on this path
  const RowPos rpos = row_pos(p, quarter * 32 + lane);

  struct Cursor {
    int f, seq, u, work, nt, img;
    uint32_t pix;
    bool valid;
  };
  auto locate = [&](Cursor& c) {  // geometry of c.work's tile for this thread
    c.valid = c.work < num_work;
    c.pix = 0xffffffffu;
    c.nt = c.img = 0;
    if (c.valid) {
      const TileCoord tc = decode_tile(p, c.work);
      const RowGeom g = row_geom(p, rpos, tc);
      c.nt = tc.nt;
      c.img = g.img;
      if (g.ok) c.pix = (uint32_t)g.pix;
    }
  };
  auto next_of = [&](const Cursor& c) {
    Cursor n = c;
    n.f += wpq;
    n.u = n.f & (nu - 1);
    const int seq = n.f >> nu_shift;
    if (seq != n.seq) {
      n.seq = seq;
      n.work = bid + seq * nblk;
      locate(n);
    }
    return n;
  };
  // Tile `seq` of this CTA lives in accumulator stage seq & 1, phase (seq >> 1) & 1.  A warp waits ONLY for tiles it has a
  // unit in.  A parity wait cannot tell a phase from the one two completions away, so the stage's completion count must
  // be pinned to {seq / 2, seq / 2 + 1} when the warp looks: not more, because tile seq + 2 cannot complete before this
  // warp has read its unit of tile seq (waiting on a tile WITHOUT a unit in it has no such bound -- it deadlocked once
  // the MMA warp was two tiles ahead); not less, because the warp's previous unit was in tile seq - 1 or seq - 2 (a
  // step of wpq <= 3 units with nu >= 2 units per tile), whose completion implies that of tile seq - 2.
  int waited = -1;

  // One unit: `in` was loaded earlier (by the previous step, or before the loop); the loads of the NEXT unit go into
  // `in_next`.  The two register sets alternate between calls, so no load result is ever copied: a register move of a
  // prefetched value would stall the warp until that load has landed (it did: ~1 us per tile, profiles/r2q_*).
  auto step = [&](const Cursor& c, CellUnitIn& in, CellUnitIn& in_next) -> Cursor {
    const Cursor n = next_of(c);
    const int col0 = c.nt * p.BN + 16 * c.u;
    const int chg0 = col0 >> 2;
#ifdef RSIS_DEBUG_TIMING
    if (n.valid) cell_unit_load(p, in_next, (p.dbg_skip & 1) ? 0xffffffffu : n.pix, ((n.nt * p.BN) >> 2) + 4 * n.u);
#else
    if (n.valid) cell_unit_load(p, in_next, n.pix, ((n.nt * p.BN) >> 2) + 4 * n.u);
#endif
    // the unit's folded affine: issued before the accumulator wait, so its (global, L1 / L2) latency hides behind that
    // wait and the tensor-memory loads instead of sitting between them and the gate math (ncu source view of the level-4
    // launch: as many stall samples on the first FFMA below as on the accumulator wait itself)
    float4 sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sc[j] = __ldg(reinterpret_cast<const float4*>(p.scale + col0 + 4 * j));
      sh[j] = __ldg(reinterpret_cast<const float4*>(p.shift + col0 + 4 * j));
    }
    if (waited != c.seq) {
      if (threadIdx.x == 0) STAMP_T(7, c.seq);
      waited = c.seq;
      mbar_wait(tfull0 + 8 * (waited & 1), (uint32_t)(waited >> 1) & 1u);
      tc_fence_after();
      if (threadIdx.x == 0) STAMP_T(3, c.seq);
    }
    const int acc = c.seq & 1;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * kStageCols;
    uint32_t r[16];
    tmem_ld16(taddr + 16 * c.u, r);
    if (p.stacked) {
      uint32_t r2[16];
      tmem_ld16(taddr + p.BN + 16 * c.u, r2);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 16; ++e) r[e] = __float_as_uint(__uint_as_float(r[e]) + __uint_as_float(r2[e]));
    } else {
      tmem_ld_wait();
    }
    // this unit's columns are in registers: one arrival per thread and unit (the stage is free for the MMAs of the tile
    // after next once all nu * 128 of them are in)
    tc_fence_before();
    mbar_arrive(tempty0 + 8 * acc);
    if (threadIdx.x == 0) STAMP_T(4, c.seq);
#ifdef RSIS_DEBUG_TIMING
    const bool ok = c.pix != 0xffffffffu && chg0 < Ch && !(p.dbg_skip & 2);
#else
    const bool ok = c.pix != 0xffffffffu && chg0 < Ch;
#endif
    float cv[4], hv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gi = fast_sigmoid(fmaf(__uint_as_float(r[4 * j + 0]), sc[j].x, sh[j].x) + in.pre[j].x);
      const float gf = fast_sigmoid(fmaf(__uint_as_float(r[4 * j + 1]), sc[j].y, sh[j].y) + in.pre[j].y);
      const float go = fast_sigmoid(fmaf(__uint_as_float(r[4 * j + 2]), sc[j].z, sh[j].z) + in.pre[j].z);
      const float gg = fast_tanh(fmaf(__uint_as_float(r[4 * j + 3]), sc[j].w, sh[j].w) + in.pre[j].w);
      const float cpj = j == 0 ? in.cp.x : (j == 1 ? in.cp.y : (j == 2 ? in.cp.z : in.cp.w));
      cv[j] = fmaf(gf, cpj, gi * gg);
      hv[j] = go * fast_tanh(cv[j]);
    }
    if (ok) {
      const size_t idx = (size_t)c.pix * Ch + chg0;
      *reinterpret_cast<float4*>(p.c_out + idx) = make_float4(cv[0], cv[1], cv[2], cv[3]);
      *reinterpret_cast<float4*>(p.h_out + idx) = make_float4(hv[0], hv[1], hv[2], hv[3]);
      if (p.h_split) store4(p.h_split, p.hs_plane, RSIS_FMT_SPLIT_BF16, (size_t)c.pix * p.hs_cs + chg0, hv);
    }
    if (p.side_max && chg0 < Ch) {  // the global nn.MaxPool2d of model.py:143 as order-preserving keys
      if (warp_one_image) {
        const int img0 = __shfl_sync(0xffffffffu, c.img, 0);
        uint32_t mine = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t m = __reduce_max_sync(0xffffffffu, ok ? float_to_key(hv[j]) : 0u);
          if (lane == j) mine = m;
        }
        if (lane < 4 && mine != 0u)
          atomicMax(p.side_max + (size_t)img0 * p.side_stride + p.side_offset + chg0 + lane, mine);
      } else if (ok) {  // maps smaller than a warp's 32 rows: rows of several images share the warp
#pragma unroll
        for (int j = 0; j < 4; ++j)
          atomicMax(p.side_max + (size_t)c.img * p.side_stride + p.side_offset + chg0 + j, float_to_key(hv[j]));
      }
    }
    if (threadIdx.x == 0 && c.u + wpq >= nu) STAMP_T(6, c.seq);
    return n;
  };

  Cursor c;
  c.f = slot;
  c.u = c.f & (nu - 1);
  c.seq = c.f >> nu_shift;
  c.work = bid + c.seq * nblk;
  locate(c);
  if (!c.valid) return;
  CellUnitIn set_a, set_b;
  cell_unit_load(p, set_a, c.pix, ((c.nt * p.BN) >> 2) + 4 * c.u);
  for (;;) {
    c = step(c, set_a, set_b);
    if (!c.valid) break;
    c = step(c, set_b, set_a);
    if (!c.valid) break;
  }
}

// The whole CTA program.  `bid` / `nblk`: this CTA's index among the `nblk` CTAs that share the problem `p` (the whole
// grid for a plain launch; a contiguous CTA range of a grouped launch, cell_group_kernel).  PW = 0: the epilogue piece
// width is taken from p.pw at run time (grouped launches mix levels that want 16 and 32).
template <bool CELL, int PW, bool SPLIT, int EW = kEpiWarps, bool CS = false>  // EW epilogue warps (then A producer, MMA issuer,
                                                                                // B producer); CS: cluster split-K
__device__ __forceinline__ void umma_cta(const UmmaMaps& maps, const UmmaParams& p, const int bid, const int nblk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 2 * kAccStages + 4];
  __shared__ uint32_t tmem_slot;

  if (threadIdx.x == 0) STAMP(0);
  // SWIZZLE_128B operand tiles need 1024-byte alignment
  const uint32_t smem_a = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_b = smem_a + p.a_stages * p.a_stage_bytes;
  const uint32_t smem_p = smem_b + p.b_stages * p.b_stage_bytes;  // pre_tma: two stages of the hoisted gate share
  const uint32_t pfull0 = smem_u32(&bars[4 * kMaxStages + 2 * kAccStages]);
  const uint32_t pempty0 = smem_u32(&bars[4 * kMaxStages + 2 * kAccStages + 2]);
  float* stage_base = reinterpret_cast<float*>(smem_raw + (smem_a - smem_u32(smem_raw)) + p.a_stages * p.a_stage_bytes +
                                               p.b_stages * p.b_stage_bytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t afull0 = smem_u32(&bars[0]);
  const uint32_t aempty0 = smem_u32(&bars[kMaxStages]);
  const uint32_t bfull0 = smem_u32(&bars[2 * kMaxStages]);
  const uint32_t bempty0 = smem_u32(&bars[3 * kMaxStages]);
  const uint32_t tfull0 = smem_u32(&bars[4 * kMaxStages]);
  const uint32_t tempty0 = smem_u32(&bars[4 * kMaxStages + kAccStages]);

  if (warp == EW && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    if (p.a_sw64 == 2) prefetch_tmap(&maps.a[1]);
    if (p.stride == 2) {
      prefetch_tmap(&maps.a[1]);
      prefetch_tmap(&maps.a[2]);
      prefetch_tmap(&maps.a[3]);
    }
    prefetch_tmap(&maps.b);
    if (CELL && p.pre_tma) prefetch_tmap(&maps.pre);
  }
  if (warp == EW + 1 && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(afull0 + 8 * s, 1);
      mbar_init(aempty0 + 8 * s, 1);
      mbar_init(bfull0 + 8 * s, 1);
      mbar_init(bempty0 + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(pfull0 + 8 * s, 1);
      mbar_init(pempty0 + 8 * s, kEpiThreads);
    }
    for (int a = 0; a < kAccStages; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      // row-wise cell epilogue: one arrival per thread and 16-column unit (BN / 16 units x 128 rows); staged epilogues:
      // one per epilogue thread and tile
      mbar_init(tempty0 + 8 * a, (CELL && !SPLIT && p.cell_rows) ? (uint32_t)p.BN * 8u : (uint32_t)kEpiThreads);
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
  if constexpr (CS) cluster_sync_all();  // the partner's barriers must be initialised before anything arrives on them
  const uint32_t tmem_base = tmem_slot;
  // PDL: everything above (barriers, TMEM, tensor-map prefetch) overlapped the previous kernel's tail, and our own
  // dependents may start their prologue right away.  Each role waits for the previous kernel (griddepcontrol.wait) right
  // before ITS first access to anything that kernel may have produced -- not here: the launch's CTAs are typically
  // resident ~10 us before the previous grid ends (in-situ trace, profiles/r2be_pass_trace.txt), and the first tile's
  // decode (cold parameter / instruction fetches, ~2 us), and with static weights the first weight boxes, fit in there.
  // The MMA issuer reads shared / tensor memory only and never waits.
  pdl_trigger();

  // The K loop walks (tap, chunk) K blocks: HALO chunk-major (an activation box = one chunk's halo tile, nine taps),
  // TAP tap-major (an activation box = `ag` chunks of one tap).  A "items" per tile: HALO -> one per chunk; TAP -> one per
  // (tap, chunk group).  A work unit is (tile, K slice): slice ks of `ksplit` covers items [ks * a_items / ksplit,
  // (ks + 1) * ...) (ag = bg = 1 whenever ksplit > 1).  Weight boxes carry `bg` consecutive K blocks of the walk and
  // never straddle a tap (TAP) / a chunk (HALO).
  const int agroups = p.agroups;
  const int a_items = p.halo ? p.chunks : p.taps * agroups;
  const int num_work = p.num_tiles * p.ksplit;

  if (warp == EW) {
    // =============================== TMA producer: activations (A ring) ===============================
    int as = 0;
    uint32_t aph = 0;
    int ps = 0;
    uint32_t pph = 0;
    if (lane == 0) STAMP(12);
    for (int work = bid; work < num_work; work += nblk) {
      int tile, ks;
      decode_work(p, work, tile, ks);
      if (lane == 0 && work == bid) STAMP_T(5, 0);
      const int item0 = (int)fdiv((uint32_t)(ks * a_items), p.dks), item1 = (int)fdiv((uint32_t)((ks + 1) * a_items), p.dks);
      if (lane == 0 && work == bid) STAMP_T(5, 1);
      const TileCoord tc = decode_tile(p, tile);
      const int tw = tc.tw, th = tc.th, ti = tc.ti;
      const int w0 = tw * p.BW, h0 = th * p.BH, i0 = ti * p.BI;
      if (lane == 0 && work == bid) STAMP_T(5, 2);
      if (work == bid) pdl_wait();  // first activation read of this CTA
      if (CELL && p.pre_tma) {
        // the tile's share of the hoisted gates: BN/32 boxes {32 columns, BW, BH, BI} -> [128 rows][128 bytes] each
        mbar_wait(pempty0 + 8 * ps, pph ^ 1u);
        if (elect_one()) {
          const int nt = tc.nt;
          mbar_arrive_expect_tx(pfull0 + 8 * ps, (uint32_t)p.p_stage_bytes);
          for (int j = 0; j < (p.BN >> 5); ++j)
            tma_load_4d(smem_p + ps * p.p_stage_bytes + j * 16384, &maps.pre, pfull0 + 8 * ps, nt * p.BN + 32 * j, w0, h0,
                        i0);
        }
        __syncwarp();
        if (++ps == 2) {
          ps = 0;
          pph ^= 1u;
        }
      }
      // (tap, chunk group) of the items walked incrementally: one division per K slice instead of three per item
      int cc = item0, kh = 0, kw = 0;
      if (!p.halo && item0 != 0) {  // (only split-K slices start inside the walk)
        const int tap0 = item0 / agroups;
        cc = item0 - tap0 * agroups;
        kh = tap0 / p.ksize;
        kw = tap0 - kh * p.ksize;
      }
      for (int ai = item0; ai < item1; ++ai) {
        if (lane == 0 && ai == item0 && work == bid) STAMP(13);
        mbar_wait(aempty0 + 8 * as, aph ^ 1u);
        if (lane == 0 && ai == item0 && work == bid) STAMP(14);
        if (elect_one()) {
          const uint32_t sa = smem_a + as * p.a_stage_bytes;
          const uint32_t bar = afull0 + 8 * as;
          mbar_arrive_expect_tx(bar, p.a_tx_bytes);
          if (ai == item0 && work == bid) STAMP(15);
          if (p.halo) {
            tma_load_5d(sa, &maps.a[0], bar, cc * kBK, w0 - 1, h0 - 1, i0, 0);
            if (p.a_sw64 == 2) tma_load_5d(sa + p.a_split_off, &maps.a[1], bar, 32, w0 - 1, h0 - 1, i0, 0);
          } else if (p.a_flat) {
            // 1x1, stride 1, the tile is 128 consecutive pixels: `ag` chunks x both planes in one box
            tma_load_4d(sa, &maps.a[0], bar, 0, (i0 * p.Ho + h0) * p.Wo + w0, 0, cc * p.ag);
          } else {
            if (p.stride == 1) {
              tma_load_5d(sa, &maps.a[0], bar, cc * kBK, w0 + kw - p.pad, h0 + kh - p.pad, i0, 0);
            } else {
              // input pixel = 2*out + k - pad: parity (k - pad) & 1, sub-grid index out + floor((k - pad) / 2)
              const int dh = kh - p.pad, dw = kw - p.pad;
              const int ph = dh & 1, pw = dw & 1;
              tma_load_5d(sa, &maps.a[ph * 2 + pw], bar, cc * kBK, w0 + ((dw - pw) >> 1), h0 + ((dh - ph) >> 1), i0, 0);
            }
          }
        }
        __syncwarp();
        if (lane == 0 && ai == item0) STAMP(2);
        if (lane == 0 && ai == item0) STAMP_T(0, (work - bid) / nblk);
        if (lane == 0 && work == bid) STAMP_T(11, ai - item0);
        if (++cc == agroups && !p.halo) {
          cc = 0;
          if (++kw == p.ksize) {
            kw = 0;
            ++kh;
          }
        }
        if (++as == p.a_stages) {
          as = 0;
          aph ^= 1u;
        }
      }
      if (lane == 0) STAMP(3);
    }
  } else if (warp == EW + 2) {
    // =============================== TMA producer: weights (B ring) ===============================
    // Its own warp, so that activation tiles run a_stages items ahead no matter how far the weight ring is.
    int bs = 0;
    uint32_t bph = 0;
    if (!p.early_b) pdl_wait();  // (static weights: nothing the previous kernel wrote is read here)
    for (int work = bid; work < num_work; work += nblk) {
      if (p.b_resident && work != bid) break;  // resident weights: loaded for the first tile only
      int tile, ks;
      decode_work(p, work, tile, ks);
      const int item0 = (int)fdiv((uint32_t)(ks * a_items), p.dks), item1 = (int)fdiv((uint32_t)((ks + 1) * a_items), p.dks);
      const int nt = tile - (int)fdiv((uint32_t)tile, p.dn) * p.tiles_n;
      // weight boxes of this work unit, in the order the MMA warp consumes them.  maps.b is {64, cout, plane, tap, chunk}.
      // HALO: per chunk, taps in boxes of bg.  TAP: per tap, chunks in boxes of bg (the last box of a tap may reach past
      // the last chunk: TMA zero-fills it and still counts the whole box).
      int o0, o1, i_first, i_last;  // outer range [o0, o1], first inner of o0, last inner (exclusive) of o1
#ifdef RSIS_DEBUG_TIMING
      int nbox = 0;
#endif
      const int inner = p.halo ? p.taps : p.chunks;
      if (p.halo) {
        o0 = item0; o1 = item1 - 1; i_first = 0; i_last = inner;
      } else {
        if (p.ksplit == 1) {  // the whole walk
          o0 = 0; i_first = 0; o1 = p.taps - 1; i_last = inner;
        } else {
          o0 = item0 / agroups; i_first = (item0 - o0 * agroups) * p.ag;
          const int last = item1 - 1;
          o1 = last / agroups;
          i_last = (last - o1 * agroups + 1) * p.ag;
          if (i_last > inner) i_last = inner;
        }
      }
      for (int o = o0; o <= o1; ++o) {
        const int ib = o == o0 ? i_first : 0, ie = o == o1 ? i_last : inner;
        for (int i = ib; i < ie; i += p.bg) {
          mbar_wait(bempty0 + 8 * bs, bph ^ 1u);
          if (elect_one()) {
            const uint32_t bar = bfull0 + 8 * bs;
            mbar_arrive_expect_tx(bar, p.b_tx_bytes);
            if (p.b_map3d)  // bg == 1: the plain {k, cout, plane} view
              tma_load_3d(smem_b + bs * p.b_stage_bytes, &maps.b, bar, ((p.halo ? i : o) * p.chunks + (p.halo ? o : i)) * kBK,
                          nt * p.BN, 0);
            else
              tma_load_5d(smem_b + bs * p.b_stage_bytes, &maps.b, bar, 0, nt * p.BN, 0, p.halo ? i : o, p.halo ? o : i);
          }
          __syncwarp();
          if (lane == 0 && work == bid) STAMP_T(10, nbox++);
          if (++bs == p.b_stages) {
            bs = 0;
            bph ^= 1u;
          }
        }
      }
    }
  } else if (warp == EW + 1) {
    // =============================== MMA issuer ===============================
    // kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128.
    // STACKED (BN <= 128): the weight stage holds [W_hi (BN rows) | W_lo (BN rows)] contiguously, so ONE MMA with
    // N = 2*BN multiplies an activation plane by both weight planes (columns [0,BN) and [BN,2BN) of the
    // accumulator, added in the epilogue): 2 MMAs per K step (x_hi, x_lo) give all four partial products.
    // ONE elected thread runs the whole issue loop, waits included (measured faster than a warp-wide loop with an
    // elected region per tap: no per-tap reconvergence).  This warp shares its scheduler with two epilogue warps, so
    // every dependent instruction between two UTCHMMAs delays the tensor pipe: resident HALO weights (the narrow,
    // many-tile decoder levels) take the unrolled straight-line form of issue_halo_resident.
    if (elect_one()) {
      const uint64_t adesc0 = make_smem_desc(smem_a, p.a_sbo, p.a_sw64 ? 4u : 2u);
      const uint64_t bdesc0 = make_smem_desc(smem_b, 1024);
      const uint32_t a_row16 = p.a_sw64 ? 4u : 8u;  // descriptor units per activation row (HALO tap shifts)
      const uint32_t n_mma = (uint32_t)(p.stacked ? 2 * p.BN : p.BN);
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((n_mma >> 3) << 17) | ((kBM >> 4) << 24);
      // Stacked weight planes: X_hi multiplies [W_hi | W_lo] (N = 2 BN), but X_lo only needs W_hi -- the first BN columns of
      // the same tile and of the same accumulator -- so its MMA is issued with N = BN: no lo x lo product, half its
      // weight-operand read and (SS-mode cost max(N/2, 32 + N/4)) 8-50 % less tensor-pipe time for it.
      const uint32_t idesc_lo = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)p.BN >> 3) << 17) | ((kBM >> 4) << 24);
      const uint32_t a_stage16 = (uint32_t)p.a_stage_bytes >> 4, b_stage16 = (uint32_t)p.b_stage_bytes >> 4;
      const uint32_t a_chunk16 = (uint32_t)p.a_chunk_bytes >> 4, b_chunk16 = (uint32_t)p.b_chunk_bytes >> 4;
      const int bg = p.bg, ag = p.ag;
      const uint32_t a_plane16 = (uint32_t)p.a_plane_bytes >> 4, b_plane16 = (uint32_t)(p.BN * 128) >> 4;
      const bool stacked = p.stacked != 0, halo = p.halo != 0, resident = p.b_resident != 0, single = p.single != 0;
      // (cells only: the unrolled form is ~5000 instructions, and the convolution kernels are already larger than the
      // instruction cache likes -- every role of the CTA runs different code)
      // Stacked weight planes only: the narrow levels this path exists for have BN <= 64.
      const bool fast = CELL && halo && resident && p.chunks == 1 && p.taps == 9 && stacked && !single;
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int work = bid; work < num_work; work += nblk) {
        int tile, ks;
        decode_work(p, work, tile, ks);
        const int item0 = (int)fdiv((uint32_t)(ks * a_items), p.dks), item1 = (int)fdiv((uint32_t)((ks + 1) * a_items), p.dks);
        const bool first_work = work == bid;
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * kStageCols;
        uint32_t accumulate = 0;
        int cc = (halo || item0 == 0) ? item0 : item0 % agroups;  // HALO: the item's chunk; TAP: its chunk group within the tap
        int bgi = 0, b_sub = 0;                   // weight box of this work unit (= its slot when resident), K block inside it
        for (int ai = item0; ai < item1; ++ai) {
          const int chunk0 = halo ? cc : cc * ag;
          const int n_in = halo ? p.taps : (p.chunks - chunk0 < ag ? p.chunks - chunk0 : ag);  // K blocks of this item
          int ksteps = (chunk0 == p.chunks - 1) ? p.last_ksteps : kBK / 16;
          if (++cc == agroups) cc = 0;  // (only read above: the NEXT item)
          mbar_wait(afull0 + 8 * as, aph);
          tc_fence_after();
          if (ai == item0) STAMP(4);
          if (ai == item0) STAMP_T(1, (work - bid) / nblk);
          if (first_work) STAMP_T(8, ai - item0);
          const uint64_t a_hi0 = adesc0 + (uint64_t)((uint32_t)as * a_stage16);
          const uint64_t a_lo0 = a_hi0 + (uint64_t)a_plane16;
          if (CELL && fast) {
            const bool wb = first_work;  // the nine taps are loaded once, for this CTA's first tile
#define RSIS_ISSUE_HALO(PASSES)                                                                                          \
  switch (ksteps) {                                                                                                      \
    case 1: issue_halo_resident<1, PASSES>(d, a_hi0, a_lo0, bdesc0, b_chunk16, b_plane16, idesc, idesc_lo, accumulate, wb, bfull0, bg); break; \
    case 2: issue_halo_resident<2, PASSES>(d, a_hi0, a_lo0, bdesc0, b_chunk16, b_plane16, idesc, idesc_lo, accumulate, wb, bfull0, bg); break; \
    case 3: issue_halo_resident<3, PASSES>(d, a_hi0, a_lo0, bdesc0, b_chunk16, b_plane16, idesc, idesc_lo, accumulate, wb, bfull0, bg); break; \
    default: issue_halo_resident<4, PASSES>(d, a_hi0, a_lo0, bdesc0, b_chunk16, b_plane16, idesc, idesc_lo, accumulate, wb, bfull0, bg); break; \
  }
#ifdef RSIS_DEBUG_TIMING
            if (p.dbg_skip & 4) {  // (diagnostic: no MMAs at all -- what the operand pipeline alone sustains)
            } else
#endif
            if (p.a_sw64 == 2) {  // 32 + 16 channels: three K steps
              const uint32_t part32 = smem_a + (uint32_t)as * (uint32_t)p.a_stage_bytes + (uint32_t)p.a_split_off;
              const uint64_t a32_hi = make_smem_desc(part32, (uint32_t)(kHaloBW + 2) * 32u, 6u);
              const uint64_t a32_lo = make_smem_desc(part32 + (uint32_t)kHaloRows * 32u, (uint32_t)(kHaloBW + 2) * 32u, 6u);
              issue_halo_resident_split48(d, a_hi0, a_lo0, a32_hi, a32_lo, bdesc0, b_chunk16, idesc, idesc_lo, accumulate, wb, bfull0,
                                          bg);
            } else if (p.a_sw64) {  // <= 32 channels: one or two K steps
              if (ksteps == 1)
                issue_halo_resident<1, 2, 4>(d, a_hi0, a_lo0, bdesc0, b_chunk16, b_plane16, idesc, idesc_lo, accumulate, wb, bfull0, bg);
              else
                issue_halo_resident<2, 2, 4>(d, a_hi0, a_lo0, bdesc0, b_chunk16, b_plane16, idesc, idesc_lo, accumulate, wb, bfull0, bg);
            } else {
              RSIS_ISSUE_HALO(2)
            }
#undef RSIS_ISSUE_HALO
            accumulate = 1u;
          } else if (!SPLIT && halo && bg == 3 && !resident && !single && ksteps == kBK / 16) {
            if (stacked)
              issue_halo_stream3<2>(d, a_hi0, a_lo0, bdesc0, b_stage16, b_chunk16, b_plane16, idesc, idesc_lo, accumulate, bfull0,
                                    bempty0, bs, bph, p.b_stages);
            else
              issue_halo_stream3<3>(d, a_hi0, a_lo0, bdesc0, b_stage16, b_chunk16, b_plane16, idesc, idesc_lo, accumulate, bfull0,
                                    bempty0, bs, bph, p.b_stages);
            bgi += 3;
            accumulate = 1u;
          } else {
            for (int bi = 0; bi < n_in; ++bi) {
              if (!halo && bi > 0) ksteps = (chunk0 + bi == p.chunks - 1) ? p.last_ksteps : kBK / 16;
              if (b_sub == 0) {  // first K block of a weight box
                if (resident) {
                  bs = bgi;  // slot = box index; its barrier completed phase 0 once and for all
                  bph = 0;
                }
                if (!(resident && !first_work)) {
                  mbar_wait(bfull0 + 8 * bs, bph);
                  tc_fence_after();
                }
                if (first_work) STAMP_T(9, bgi);
              }
              if (ai == item0 && bi == 0) STAMP(5);
              // HALO: tap (kh, kw) = the tile shifted by kh*10 + kw rows of 128 bytes (8 descriptor units per row);
              // TAP: chunk bi of the activation box
              const uint64_t toff = halo ? (uint64_t)((uint32_t)(bi + 7 * ((bi * 11) >> 5)) * a_row16)
                                         : (uint64_t)((uint32_t)bi * a_chunk16);
              const uint64_t a_hi = a_hi0 + toff, a_lo = a_lo0 + toff;
              const uint64_t b_hi = bdesc0 + (uint64_t)((uint32_t)bs * b_stage16 + (uint32_t)b_sub * b_chunk16);
              // the common full block (four K steps, split operands) straight-line with compile-time descriptor offsets;
              // partial last chunks and the single-pass bf16 mode take a compact loop (code size: see above)
              if (ksteps == kBK / 16 && !single) {
                if (stacked)
                  issue_kblock<kBK / 16, 2>(d, a_hi, a_lo, b_hi, b_plane16, idesc, idesc_lo, accumulate);
                else
                  issue_kblock<kBK / 16, 3>(d, a_hi, a_lo, b_hi, b_plane16, idesc, idesc_lo, accumulate);
              } else {
                for (int k = 0; k < ksteps; ++k) {
                  const uint64_t adv = (uint64_t)(2 * k);
                  umma_bf16(d, a_hi + adv, b_hi + adv, idesc, k == 0 ? accumulate : 1u);
                  if (!single) {
                    if (!stacked) umma_bf16(d, a_hi + adv, b_hi + b_plane16 + adv, idesc, 1u);
                    umma_bf16(d, a_lo + adv, b_hi + adv, idesc_lo, 1u);
                  }
                }
              }
              accumulate = 1u;
              // last K block of the weight box: boxes end with the tap (TAP) / the chunk (HALO)
              if (++b_sub == bg || (halo ? bi == n_in - 1 : chunk0 + bi == p.chunks - 1)) {
                b_sub = 0;
                ++bgi;
                if (!resident) {
                  umma_commit(bempty0 + 8 * bs);  // weight slot free once these MMAs have read it
                  if (++bs == p.b_stages) {
                    bs = 0;
                    bph ^= 1u;
                  }
                }
              }
            }
          }
          umma_commit(aempty0 + 8 * as);  // activation slot free
          if (++as == p.a_stages) {
            as = 0;
            aph ^= 1u;
          }
        }
        umma_commit(tfull0 + 8 * acc);  // accumulator complete
        STAMP(6);
        STAMP_T(2, (work - bid) / nblk);
        if (++acc == kAccStages) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp < EW) {
    // =============================== epilogue (warps 0-7) ===============================
    float* stage = stage_base + warp * kStageFloats;
    pdl_wait();  // the epilogues prefetch residual / state operands right away
    if (threadIdx.x == 0) STAMP(1);
    if (CELL && !SPLIT && p.cell_rows) {
      cell_rows_epilogue(p, tmem_base, tfull0, tempty0, warp, lane, bid, nblk, EW / 4);
    } else if constexpr (PW != 0) {
      epilogue_role<CELL, PW, SPLIT, CS>(p, tmem_base, tfull0, tempty0, stage, warp, lane, bid, nblk, smem_a + (uint32_t)p.x_off,
                                     pempty0);  // (pempty0: 256 arrivals, otherwise unused -- the exchange barrier)
    }
    // PW == 0 (grouped launch): row-wise epilogue only -- convlstm_cell_group_umma refuses to run without it, and
    // leaving the two staged-transpose variants out keeps ~4000 instructions out of the kernel
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) STAMP(11);
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// One instantiation per (epilogue kind, piece width, split-K or not): each carries a single epilogue variant, which
// keeps the 11 differently-specialised warps of a CTA inside the instruction cache.
template <bool CELL, int PW, bool SPLIT, bool CS = false>
__global__ void __launch_bounds__(kThreadsUmma, 1)
conv_umma_kernel(const __grid_constant__ UmmaMaps maps, const UmmaParams p) {
#ifdef RSIS_DEBUG_TIMING
  trace_begin(p);
#endif
  umma_cta<CELL, PW, SPLIT, kEpiWarps, CS>(maps, p, (int)blockIdx.x, (int)gridDim.x);
#ifdef RSIS_DEBUG_TIMING
  trace_end(p);
#endif
}

// Grouped launch: up to kMaxGroup independent ConvLSTM cells (different decoder levels, different time-steps: the
// anti-diagonal of the (level, step) wavefront, model.py:129-165) in ONE launch.  Problem i owns the CTA range
// [first[i], first[i+1]); every CTA runs the ordinary persistent program on its own problem.  No inter-CTA dependency
// exists inside the launch, so the fixed costs of a launch (tensor-map fetch, first cold TMA, epilogue drain) are paid
// once per wavefront instead of once per level, and the small levels run beside the large ones.
constexpr int kMaxGroup = 5;
struct alignas(64) CellGroup {
  UmmaMaps maps[kMaxGroup];
  UmmaParams p[kMaxGroup];
  int first[kMaxGroup + 1];
  int n;
};

#ifndef RSIS_GROUP_EPI_WARPS
#define RSIS_GROUP_EPI_WARPS 12
#endif
constexpr int kGroupEpiWarps = RSIS_GROUP_EPI_WARPS;          // three epilogue warps per TMEM lane quarter
constexpr int kThreadsGroup = kGroupEpiWarps * 32 + 128;       // + A producer, MMA issuer, B producer, one idle warp
__global__ void __launch_bounds__(kThreadsGroup, 1) cell_group_kernel(const __grid_constant__ CellGroup g) {
  int i = 0;
#pragma unroll
  for (int k = 1; k < kMaxGroup; ++k)
    if (k < g.n && (int)blockIdx.x >= g.first[k]) i = k;
#ifdef RSIS_DEBUG_TIMING
  trace_begin(g.p[0]);
#endif
  umma_cta<true, 0, false, kGroupEpiWarps>(g.maps[i], g.p[i], (int)blockIdx.x - g.first[i], g.first[i + 1] - g.first[i]);
#ifdef RSIS_DEBUG_TIMING
  trace_end(g.p[0]);
  // per-cell end time (max over its CTAs) and launch start (min over all CTAs): slots 200 + i and 199 of the stamp table
  if (threadIdx.x == 0 && g.p[0].counters) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    atomicMax(reinterpret_cast<unsigned long long*>(g.p[0].counters) + 256 + 200 + i, t);
  }
#endif
}

// ===================================================================================================================
// Swapped-operand ConvLSTM cell for the narrow, high-resolution decoder levels (4*Ch = 32 or 64 gate columns, at most
// 64 input channels once the skip share is hoisted; levels 3-4 of the reference decoder).
//
// An SS-mode tcgen05.mma costs about the same whatever its N, so 128 pixels x 64 gate columns per instruction leaves
// the tensor pipe mostly idle.  Here the roles are swapped: D[gate rows][256 pixels] = W[gate rows][K] * X[pixels][K]^T.
//   A = weights of one tap: rows (gate column r, plane) -> 2r + plane, i.e. hi and lo bf16 planes of a gate column in
//       ADJACENT accumulator lanes (64 or 128 rows; one TMA box {64 k, 2 planes, 4*Ch} of the same packed tensor);
//   B = activations: the 8 x 32 output tile's 10 x 34 input halo, staged ONCE per tile (one TMA box); tap (kh, kw) is
//       the descriptor shifted by kh*10 + kw rows, N = 256 = 32 core-matrix groups 1280 bytes apart;
//   two MMAs per K step (x_hi, x_lo): lanes 2r / 2r+1 accumulate (W_hi + ...)(x_hi + x_lo) and (W_lo ...)(...), added
//   with one lane shuffle in the epilogue -> all four partial products for half the instructions per pixel.
// Epilogue: accumulator lanes are gate columns and columns are pixels; a warp owns 4 hidden channels x 128 pixels,
// transposes 32-pixel chunks through shared memory and finishes with lane = (pixel, channel).
constexpr int kSwTileH = 32;
constexpr int kSwHaloRows = (kHaloBW + 2) * (kSwTileH + 2);  // 340
constexpr int kSwActPlaneBytes = kSwHaloRows * 128;           // 43520
constexpr int kSwActBytes = 2 * kSwActPlaneBytes;             // 87040 = 85 KB (1024-byte multiple)

__device__ __forceinline__ size_t swap_pix(const UmmaParams& p, int img, int h0, int w0, int n) {
  return ((size_t)img * p.Ho + h0 + (n >> 3)) * p.Wo + w0 + (n & 7);
}
// Epilogue of the swapped cell.  warp = (lane quarter q, pixel half).  The accumulator lanes of quarter q are hidden
// channels 4q..4q+3.  With 64 gate columns all four quarters carry data (PAIR = false); with 32 only quarters 0-1 do,
// so the warps of quarters 2-3 (which cannot read those TMEM lanes) pair up with warp q-2 (PAIR = true): the reader
// stages a 32-pixel chunk in shared memory and both warps finish 16 pixels of it.
template <bool PAIR>
__device__ __forceinline__ void swap_epilogue(const UmmaParams& p, uint32_t tmem_base, uint32_t tfull0,
                                              uint32_t tempty0, float* stage_base, int warp, int lane) {
  constexpr int NIT = PAIR ? 2 : 4;  // 8-pixel iterations of a 32-pixel chunk this warp finishes
  const int q = warp & 3, half = warp >> 2;
  const int qr = PAIR ? (q & 1) : q;
  const bool reader = !PAIR || q < 2;
  const int it0 = (PAIR && q >= 2) ? 2 : 0;
  const int slot = PAIR ? half * 2 + qr : warp;
  const int Ch = p.Cout >> 2;
  const int c = lane & 3;
  const int chg = 4 * qr + c;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  float* stage = stage_base + slot * kStageFloats;
  const float4 sc = __ldg(reinterpret_cast<const float4*>(p.scale + 4 * chg));
  const float4 sh = __ldg(reinterpret_cast<const float4*>(p.shift + 4 * chg));
  struct Pre {
    float cp[NIT];
    float4 pre[NIT];
  };
  auto pair_sync = [&]() {
    if constexpr (PAIR) {
      // one named barrier per (reader, helper) pair: ids 2..5
      if (slot == 0) asm volatile("bar.sync 2, 64;" ::: "memory");
      else if (slot == 1) asm volatile("bar.sync 3, 64;" ::: "memory");
      else if (slot == 2) asm volatile("bar.sync 4, 64;" ::: "memory");
      else asm volatile("bar.sync 5, 64;" ::: "memory");
    } else {
      __syncwarp();
    }
  };
  auto prefetch = [&](Pre& f, int img, int h0, int w0, int col0) {
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const size_t pix = swap_pix(p, img, h0, w0, col0 + (it0 + i) * 8 + (lane >> 2));
      f.cp[i] = p.c_prev ? __ldg(p.c_prev + pix * Ch + chg) : 0.f;
      f.pre[i] = p.preact ? __ldg(reinterpret_cast<const float4*>(p.preact + (pix * Ch + chg) * 4))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  int acc = 0;
  uint32_t acc_phase = 0;
  Pre cur;
  bool first = true;
  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
    const int w0 = (r % p.tiles_w) * kHaloBW, h0 = (r / p.tiles_w) * kSwTileH;
    if (first) prefetch(cur, img, h0, w0, 128 * half);
    first = false;
    mbar_wait(tfull0 + 8 * acc, acc_phase);
    tc_fence_after();
    if (threadIdx.x == 0) STAMP_T(3, (tile - (int)blockIdx.x) / (int)gridDim.x);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kStageCols;
    uint32_t best = 0u;
    for (int j = 0; j < 4; ++j) {
      const int col0 = 128 * half + 32 * j;
      Pre nxt;
      if (j < 3) {
        prefetch(nxt, img, h0, w0, col0 + 32);
      } else if (tile + (int)gridDim.x < p.num_tiles) {
        const int t2 = tile + (int)gridDim.x;
        const int img2 = t2 / tiles_per_img, r2 = t2 - img2 * tiles_per_img;
        prefetch(nxt, img2, (r2 / p.tiles_w) * kSwTileH, (r2 % p.tiles_w) * kHaloBW, 128 * half);
      }
      if (reader) {
        uint32_t rr[32];
        tmem_ld32(taddr + col0, rr);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          // lanes 2r / 2r+1 hold the hi- / lo-weight-plane partial sums of gate column r
          const float v = __uint_as_float(rr[e]);
          stage[e * kStagePitch + lane] = v + __shfl_xor_sync(0xffffffffu, v, 1);
        }
      }
      pair_sync();
#pragma unroll
      for (int i = 0; i < NIT; ++i) {
        const int px = (it0 + i) * 8 + (lane >> 2);
        const float* g = stage + px * kStagePitch + 8 * c;
        const float gi = fast_sigmoid(fmaf(g[0], sc.x, sh.x) + cur.pre[i].x);
        const float gf = fast_sigmoid(fmaf(g[2], sc.y, sh.y) + cur.pre[i].y);
        const float go = fast_sigmoid(fmaf(g[4], sc.z, sh.z) + cur.pre[i].z);
        const float gg = fast_tanh(fmaf(g[6], sc.w, sh.w) + cur.pre[i].w);
        const float cv = fmaf(gf, cur.cp[i], gi * gg);
        const float hv = go * fast_tanh(cv);
        const size_t pix = swap_pix(p, img, h0, w0, col0 + px);
        const size_t idx = pix * Ch + chg;
        p.c_out[idx] = cv;
        p.h_out[idx] = hv;
        if (p.h_split) {
          __nv_bfloat16 hi, lo;
          split_bf16(hv, hi, lo);
          const size_t k2 = pix * p.hs_cs + chg;
          p.h_split[k2] = hi;
          p.h_split[k2 + p.hs_plane] = lo;
        }
        const uint32_t key = float_to_key(hv);
        best = key > best ? key : best;
      }
      pair_sync();
      cur = nxt;
    }
    tc_fence_before();
    mbar_arrive(tempty0 + 8 * acc);
    if (threadIdx.x == 0) STAMP_T(4, (tile - (int)blockIdx.x) / (int)gridDim.x);
    if (p.side_max) {
      // global nn.MaxPool2d (model.py:143): lanes with the same lane % 4 hold the same channel
#pragma unroll
      for (int s2 = 4; s2 < 32; s2 <<= 1) {
        const uint32_t o = __shfl_xor_sync(0xffffffffu, best, s2);
        best = o > best ? o : best;
      }
      if (lane < 4) atomicMax(p.side_max + (size_t)img * p.side_stride + p.side_offset + chg, best);
    }
    if (++acc == kAccStages) {
      acc = 0;
      acc_phase ^= 1u;
    }
  }
}

__global__ void __launch_bounds__(kThreadsUmma, 1)
cell_swap_kernel(const __grid_constant__ UmmaMaps maps, const UmmaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[4 + 2 * kMaxStages + 2 * kAccStages];
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_x = (smem_u32(smem_raw) + 1023u) & ~1023u;  // activation halo tiles (a_stages of them)
  const uint32_t smem_w = smem_x + p.a_stages * kSwActBytes;       // weight ring / resident weight set
  float* stage_base = reinterpret_cast<float*>(smem_raw + (smem_x - smem_u32(smem_raw)) + p.a_stages * kSwActBytes +
                                               p.b_stages * p.b_stage_bytes);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t xfull0 = smem_u32(&bars[0]);
  const uint32_t xempty0 = smem_u32(&bars[2]);
  const uint32_t wfull0 = smem_u32(&bars[4]);
  const uint32_t wempty0 = smem_u32(&bars[4 + kMaxStages]);
  const uint32_t tfull0 = smem_u32(&bars[4 + 2 * kMaxStages]);
  const uint32_t tempty0 = smem_u32(&bars[4 + 2 * kMaxStages + kAccStages]);

  if (warp == kEpiWarps && lane == 0) {
    prefetch_tmap(&maps.a[0]);
    prefetch_tmap(&maps.b);
  }
  if (warp == kEpiWarps + 1 && lane == 0) {
    for (int s2 = 0; s2 < 2; ++s2) {
      mbar_init(xfull0 + 8 * s2, 1);
      mbar_init(xempty0 + 8 * s2, 1);
    }
    for (int s2 = 0; s2 < kMaxStages; ++s2) {
      mbar_init(wfull0 + 8 * s2, 1);
      mbar_init(wempty0 + 8 * s2, 1);
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
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) STAMP(0);

  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int ksteps = p.last_ksteps;  // one channel chunk only

  if (warp == kEpiWarps) {
    // ---- activation producer: one halo box per tile
    uint32_t ph = 0;
    int xs = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int img = tile / tiles_per_img, r = tile - img * tiles_per_img;
      const int w0 = (r % p.tiles_w) * kHaloBW, h0 = (r / p.tiles_w) * kSwTileH;
      mbar_wait(xempty0 + 8 * xs, ph ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(xfull0 + 8 * xs, (uint32_t)kSwActBytes);
        tma_load_5d(smem_x + xs * kSwActBytes, &maps.a[0], xfull0 + 8 * xs, 0, w0 - 1, h0 - 1, img, 0);
      }
      __syncwarp();
      if (lane == 0) STAMP_T(0, (tile - (int)blockIdx.x) / (int)gridDim.x);
      if (++xs == p.a_stages) {
        xs = 0;
        ph ^= 1u;
      }
    }
  } else if (warp == kEpiWarps + 2) {
    // ---- weight producer: nine tap boxes per tile, or once per CTA when they stay resident
    int ws = 0;
    uint32_t wph = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      if (p.b_resident && tile != (int)blockIdx.x) break;
      for (int tap = 0; tap < 9; ++tap) {
        mbar_wait(wempty0 + 8 * ws, wph ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(wfull0 + 8 * ws, p.b_tx_bytes);
          tma_load_3d(smem_w + ws * p.b_stage_bytes, &maps.b, wfull0 + 8 * ws, tap * kBK, 0, 0);
        }
        __syncwarp();
        if (++ws == p.b_stages) {
          ws = 0;
          wph ^= 1u;
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ---- MMA issuer: M = 128 (weight rows; only 2 * 4*Ch of them are meaningful), N = 256 pixels
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((kBM >> 4) << 24);
    const uint64_t wdesc0 = make_smem_desc(smem_w, 1024);
    const uint64_t xdesc0 = make_smem_desc(smem_x, (kHaloBW + 2) * 128);
    int ws = 0, xs = 0;
    uint32_t wph = 0, xph = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1u);
      tc_fence_after();
      mbar_wait(xfull0 + 8 * xs, xph);
      tc_fence_after();
      if (lane == 0) STAMP_T(1, (tile - (int)blockIdx.x) / (int)gridDim.x);
      const uint32_t d = tmem_base + acc * kStageCols;
      uint32_t accumulate = 0;
      for (int tap = 0; tap < 9; ++tap) {
        if (p.b_resident) {
          ws = tap;
          wph = 0;
        }
        if (!(p.b_resident && tile != (int)blockIdx.x)) {
          mbar_wait(wfull0 + 8 * ws, wph);
          tc_fence_after();
        }
        const int kh = tap / 3, kw = tap - kh * 3;
        const uint64_t a = wdesc0 + (uint64_t)((uint32_t)(ws * p.b_stage_bytes) >> 4);
        const uint64_t b_hi = xdesc0 + (uint64_t)((uint32_t)(xs * kSwActBytes + (kh * (kHaloBW + 2) + kw) * 128) >> 4);
        const uint64_t b_lo = b_hi + (uint64_t)(kSwActPlaneBytes >> 4);
        if (elect_one()) {
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t adv = (uint64_t)(k * 32 >> 4);
            umma_bf16(d, a + adv, b_hi + adv, idesc, accumulate);
            umma_bf16(d, a + adv, b_lo + adv, idesc, 1u);
            accumulate = 1u;
          }
          if (!p.b_resident) umma_commit(wempty0 + 8 * ws);
          if (tap == 8) {
            umma_commit(xempty0 + 8 * xs);
            umma_commit(tfull0 + 8 * acc);
          }
        }
        __syncwarp();
        accumulate = 1u;
        if (!p.b_resident && ++ws == p.b_stages) {
          ws = 0;
          wph ^= 1u;
        }
      }
      if (lane == 0) STAMP_T(2, (tile - (int)blockIdx.x) / (int)gridDim.x);
      if (++xs == p.a_stages) {
        xs = 0;
        xph ^= 1u;
      }
      if (++acc == kAccStages) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else if (warp < kEpiWarps) {
    if (2 * p.Cout < 128)
      swap_epilogue<true>(p, tmem_base, tfull0, tempty0, stage_base, warp, lane);
    else
      swap_epilogue<false>(p, tmem_base, tfull0, tempty0, stage_base, warp, lane);
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
int g_early_b = 1;         // RSIS_B200_EARLY_B=0: weight boxes wait for the previous kernel even with static weights (A-B timing)
int g_max_a_stages = 4;    // RSIS_B200_ASTAGES: most activation stages beside resident weights (cold halo boxes land ~2.5 us after issue)
int g_plan_lo = 0;         // RSIS_B200_PLAN_LO=1: the planner charges the X_lo MMA of a stacked K step at N = BN
int g_csplit = 0;          // RSIS_B200_CSPLIT=1: cluster split-K plans (two-CTA clusters, DSMEM reduction) may be chosen
int g_sw64 = 1;            // RSIS_B200_SW64=0: 64-channel SWIZZLE_128B halo boxes for narrow sources too (A-B timing)
int g_split_k = 1;         // RSIS_B200_SPLITK=0 disables split-K (debug / A-B timing)
int g_force_bn = 0;        // RSIS_B200_BN forces the output-channel tile width (debug)
#ifdef RSIS_DEBUG_TIMING
unsigned long long* g_trace = nullptr;  // rsis_debug_trace: rows of 24 x u64, one per tcgen05 launch set up
int g_trace_rows = 0, g_trace_next = 0;
#endif
unsigned* g_debug_counters = nullptr;  // set by the last non-swapped setup when RSIS_B200_DEBUG_TIMING is on
int g_swap = 1;            // RSIS_B200_SWAP=0 disables the swapped-operand cell kernel for the narrow levels
int g_print_plan = 0;      // RSIS_B200_PRINT_PLAN=1 logs the tile plan of every launch to stderr
int g_pdl = 1;             // RSIS_B200_PDL=0: plain stream-ordered launches (no programmatic dependent launch)
int g_static_weights = 0;  // rsis_set_static_weights: weight boxes may be issued before griddepcontrol.wait
int g_precision = 0;       // rsis_set_precision: 0 = split bf16 (fp32-grade products), 1 = single-pass bf16 operands
int g_pre_tma = 0;         // RSIS_B200_PRE_TMA=1: the hoisted gate share of a tile is staged in shared memory by TMA (no measured
                           // gain at B=8, and its two stages cost the third activation stage: off by default)
int g_cell_rows = 1;       // RSIS_B200_CELL_ROWS=0: the staged-transpose cell epilogue instead of the row-wise one (A/B timing)
int g_mma_model = 1;       // RSIS_B200_MMA_MODEL=0: planner assumes 70 ns per MMA whatever its N (round-1 model)
int g_b_resident = 1;      // RSIS_B200_BRES=0 disables weight residency (debug / A-B timing)
int g_bmap3d = 1;          // RSIS_B200_BMAP3D=0: the 5-D weight view even for single-block boxes (A/B timing)
int g_box_group = 1;       // RSIS_B200_BOXGROUP=0: one (tap, chunk) K block per TMA box, as in round 1 (A/B timing)
std::once_flag g_once;

constexpr int kSmemLimit = 227 * 1024;
constexpr int kDynSmem = kSmemLimit - 1024;  // static barriers live beside it

template <bool CELL, int PW, bool SPLIT, bool CS = false>
cudaError_t set_smem_attr() {
  return cudaFuncSetAttribute(conv_umma_kernel<CELL, PW, SPLIT, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem);
}

void init_once() {
  if (const char* e = getenv("RSIS_B200_HALO")) g_halo_enabled = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_SW64")) g_sw64 = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_CSPLIT")) g_csplit = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_PLAN_LO")) g_plan_lo = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_ASTAGES")) g_max_a_stages = atoi(e) < 2 ? 2 : (atoi(e) > kMaxStages ? kMaxStages : atoi(e));
  if (const char* e = getenv("RSIS_B200_EARLY_B")) g_early_b = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_SPLITK")) g_split_k = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_BN")) g_force_bn = atoi(e);
  if (const char* e = getenv("RSIS_B200_BRES")) g_b_resident = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_BOXGROUP")) g_box_group = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_BMAP3D")) g_bmap3d = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_PDL")) g_pdl = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_MMA_MODEL")) g_mma_model = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_CELL_ROWS")) g_cell_rows = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_PRE_TMA")) g_pre_tma = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_PRINT_PLAN")) g_print_plan = atoi(e) != 0;
  if (const char* e = getenv("RSIS_B200_SWAP")) g_swap = atoi(e) != 0;

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
      (e = set_smem_attr<false, 32, false>()) != cudaSuccess || (e = set_smem_attr<false, 32, true>()) != cudaSuccess ||
      (e = set_smem_attr<false, 16, false>()) != cudaSuccess || (e = set_smem_attr<false, 16, true>()) != cudaSuccess ||
      (e = set_smem_attr<true, 32, false>()) != cudaSuccess || (e = set_smem_attr<true, 32, true>()) != cudaSuccess ||
      (e = set_smem_attr<true, 16, false>()) != cudaSuccess || (e = set_smem_attr<true, 16, true>()) != cudaSuccess ||
      (e = set_smem_attr<false, 32, false, true>()) != cudaSuccess || (e = set_smem_attr<false, 16, false, true>()) != cudaSuccess ||
      (e = cudaFuncSetAttribute(cell_swap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem)) !=
          cudaSuccess ||
      (e = cudaFuncSetAttribute(cell_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynSmem)) !=
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
int encode_act_map(CUtensorMap* m, const rsis_tensor& t, int sub, int ph, int pw, int BW, int BH, int BI,
                   int planes = 2, int box_c = kBK) {
  const size_t C = t.c, W = t.w, H = t.h, N = t.n, P = pitch(t);
  char* base = reinterpret_cast<char*>(t.data) + ((size_t)ph * W + pw) * P * 2;
  cuuint64_t dims[5] = {C, W / sub, H / sub, N, 2};
  cuuint64_t strides[4] = {P * 2 * sub, W * P * 2 * sub, H * W * P * 2, N * H * W * P * 2};
  cuuint32_t box[5] = {(cuuint32_t)box_c, (cuuint32_t)BW, (cuuint32_t)BH, (cuuint32_t)BI, (cuuint32_t)planes};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  // box_c = 32: rows of 64 bytes, SWIZZLE_64B; box_c = 16: rows of 32 bytes, SWIZZLE_32B (UmmaParams::a_sw64)
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        box_c == kBK ? CU_TENSOR_MAP_SWIZZLE_128B : (box_c == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RSIS_OK : RSIS_ERR_CUDA;
}

int encode_weight_map3(CUtensorMap* m, const void* w, int cout_pad, int k_pad, int BN, int planes) {
  cuuint64_t dims[3] = {(cuuint64_t)k_pad, (cuuint64_t)cout_pad, 2};
  cuuint64_t strides[2] = {(cuuint64_t)k_pad * 2, (cuuint64_t)cout_pad * k_pad * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)BN, (cuuint32_t)planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RSIS_OK : RSIS_ERR_CUDA;
}

// Packed weights [2 planes][cout_pad][tap][chunk][64] as the 5-D view {64, cout, plane, tap, chunk}: a box
// {64, BN, planes, box_taps, box_chunks} lands as [chunk][tap][plane][BN rows][128 bytes], i.e. as consecutive
// (tap, chunk) K blocks of the [W_hi | W_lo] stage layout -- several K blocks per TMA instruction.
// 4-D flat-pixel view {64, pixels, plane, chunk} of a split-bf16 NHWC activation whose channel count is a multiple of
// 64: the 128-pixel tile of a 1x1 stride-1 convolution with `box_chunks` channel chunks in one box, landing as
// [chunk][plane][128 rows][128 bytes].
int encode_act_map_flat(CUtensorMap* m, const rsis_tensor& t, int box_chunks, int planes) {
  const size_t P = pitch(t), npix = (size_t)t.n * t.h * t.w;
  cuuint64_t dims[4] = {(cuuint64_t)kBK, npix, 2, (cuuint64_t)(t.c / kBK)};
  cuuint64_t strides[3] = {P * 2, npix * P * 2, (cuuint64_t)kBK * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)kBM, (cuuint32_t)planes, (cuuint32_t)box_chunks};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.data, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RSIS_OK : RSIS_ERR_CUDA;
}

int encode_weight_map(CUtensorMap* m, const void* w, int cout_pad, int taps, int chunks, int BN, int planes, int box_taps,
                      int box_chunks) {
  const cuuint64_t k_pad = (cuuint64_t)taps * chunks * kBK;
  cuuint64_t dims[5] = {(cuuint64_t)kBK, (cuuint64_t)cout_pad, 2, (cuuint64_t)taps, (cuuint64_t)chunks};
  cuuint64_t strides[4] = {k_pad * 2, (cuuint64_t)cout_pad * k_pad * 2, (cuuint64_t)chunks * kBK * 2, (cuuint64_t)kBK * 2};
  cuuint32_t box[5] = {(cuuint32_t)kBK, (cuuint32_t)BN, (cuuint32_t)planes, (cuuint32_t)box_taps, (cuuint32_t)box_chunks};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(w), dims, strides, box, estr,
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

constexpr size_t kScratchFloats = (size_t)148 * kBM * kStageCols;   // one partial accumulator per SM
constexpr size_t kCounterBytes = 4096;
constexpr size_t kWorkspaceBytes = kCounterBytes + kScratchFloats * sizeof(float);

// Tile-shape planner.  An SS-mode tcgen05.mma costs ~128 cycles of A-operand (activation) reads whatever its N, so the
// time of a launch is (MMAs per CTA) x 128 cycles: pick the output-channel width BN (stacking the hi|lo weight planes
// along N when 2*BN <= 256) and, when a launch has fewer work units than SMs, a K split that minimises it.
struct Plan {
  int BN, stacked, ksplit, halo;
  long long cost;
  int csplit = 0;  // 1: ksplit == 2 as a two-CTA cluster with a DSMEM reduction
};

// Cost of one tcgen05.mma (M = 128, K = 16, kind::f16, SS mode) in ns, from scripts/mma_probe.cu on B200
// (profiles/r2b_mma_probe.txt): max(N/2, 32 + N/4) cycles -- the tensor pipe needs N/2 cycles, the shared-memory
// operand reads (A: 4 KB, B: N x 32 B at 128 B/clk) 32 + N/4.  The extra 8 cycles and the 1.9 GHz stand for the
// operand-port contention with TMA fills seen in the in-kernel stamps (N = 256: 70 ns measured, 67 modelled).
inline double mma_ns(int n_mma) {
  if (!g_mma_model) return 70.0;  // RSIS_B200_MMA_MODEL=0: the round-1 constant (A/B timing)
  const double a = n_mma / 2.0, b = 40.0 + n_mma / 4.0;
  return (a > b ? a : b) / 1.9;
}

thread_local int t_force_bn = 0;  // > 0: output-channel tile width for the next plan (grouped-launch tuning override)

Plan make_plan(int m_tiles_halo, int m_tiles_tap, bool halo_ok, int cout, int taps, int chunks, int last_ksteps,
               bool can_split, bool cell, int ctas, bool single = false) {
  // Cost model in nanoseconds, calibrated on B200 with in-kernel %globaltimer stamps and graph-replay timings
  // (scripts/stamp_probe.py, scripts/gap_probe.py):
  //   launch -> first MMA and exit: ~3000;  one tcgen05.mma instruction: ~70 whatever its N;
  //   TMA ingest of one SM: ~80 bytes/ns, of the whole chip (L2 -> SMs): ~12000 bytes/ns;
  //   finishing one 32 x 32 accumulator piece on one epilogue warp: ~900 (conv) / ~1300 (cell), 8 warps per CTA;
  //   split-K hand-over (park the partial, wait for the slowest slice, reduce): ~4500 + 110 per (unit, slice).
  const double kFixed = 3000, kPiece = cell ? 1300 : 900, kSmBw = 80, kChipBw = 12000;
  if (ctas <= 0 || ctas > g_num_sms) ctas = g_num_sms;  // CTAs this problem may use (its share of a grouped launch)
  const double ksteps_tile = (double)taps * ((chunks - 1) * (kBK / 16) + last_ksteps);
  const int items = taps * chunks;
  Plan best{0, 0, 1, 0, -1};
  const int all[4] = {256, 128, 64, 32};
  for (int ci = 0; ci < 4; ++ci) {
    const int BN = all[ci];
    if (BN >= 2 * cout && BN > 32) continue;              // wider than the layer: nothing but padding
    if (cout <= 128 && BN > 128) continue;
    if (g_force_bn && BN != g_force_bn) continue;
    if (t_force_bn && BN != t_force_bn) continue;
    const int stacked = (BN <= 128 && !single) ? 1 : 0;
    // MMAs per K step x their cost: stacked = X_hi x [W_hi | W_lo] (N = 2 BN) + X_lo x W_hi (N = BN); RSIS_B200_PLAN_LO=0
    // keeps the model of two N = 2 BN MMAs the plans were first calibrated with (A/B timing)
    const double mpk = single ? 1 : (stacked ? (g_plan_lo ? 1 : 2) : 3);
    const double kMma = (stacked && g_plan_lo) ? mma_ns(2 * BN) + mma_ns(BN) : mma_ns(stacked ? 2 * BN : BN);
    const int tiles_n = ceil_div(cout, BN);
    const double b_item = (single ? 1.0 : 2.0) * BN * 128;
    const double pieces_per_warp = BN >= 64 ? ceil_div(4 * (BN / 32), kEpiWarps) : 0.6;
    const double epi_tile = pieces_per_warp * kPiece;
    // (a) no split: persistent CTAs, HALO staging when eligible; MMAs of tile i+1 overlap the epilogue of tile i
    if (!(halo_ok && BN == 256)) {  // a HALO stage + 64 KB weight stages do not leave room for a pipeline
      const double tiles = (double)(halo_ok ? m_tiles_halo : m_tiles_tap) * tiles_n;
      const double per_cta = ceil(tiles / ctas);
      const double bytes_tile = (halo_ok ? chunks * 2.0 * kHaloRows * 128 : items * 32768.0) + items * b_item;
      const double mma_tile = ksteps_tile * mpk * kMma;
      double tile_t = mma_tile > epi_tile ? mma_tile : epi_tile;
      if (bytes_tile / kSmBw > tile_t) tile_t = bytes_tile / kSmBw;
      double cost = kFixed + per_cta * tile_t + epi_tile;
      const double chip = kFixed + tiles * bytes_tile / kChipBw;
      if (chip > cost) cost = chip;
      if (best.cost < 0 || cost < best.cost) best = Plan{BN, stacked, 1, halo_ok ? 1 : 0, (long long)cost};
    }
    // (c) split-K over 64-channel chunks with HALO staging (each slice still stages its halo once per chunk)
    if (can_split && g_split_k && halo_ok && BN != 256 && chunks >= 2) {
      const double tiles = (double)m_tiles_halo * tiles_n;
      int S = (int)(g_num_sms / tiles);
      if (S > chunks) S = chunks;
      if (tiles <= g_num_sms && S >= 2) {
        const double chunks_s = ceil_div(chunks, S);
        const double mma = chunks_s * taps * (kBK / 16) * mpk * kMma;
        const double load = chunks_s * (2.0 * kHaloRows * 128 + taps * b_item) / kSmBw;
        const double units_per_warp = ceil_div((kBM / 4) * ceil_div(BN, 32), kEpiWarps * S);
        double cost = kFixed + (mma > load ? mma : load) + 4500 + units_per_warp * (300 + 110.0 * S * (stacked ? 2 : 1));
        if (cost < best.cost) best = Plan{BN, stacked, S, 1, (long long)cost};
      }
    }
    // (d) cluster split-K (RSIS_B200_CSPLIT=1): the two K halves of a tile on the two CTAs of a cluster, the second half's
    // partial tile handed over through distributed shared memory (~1.2 us) instead of the L2 scratch + counters of (b) /
    // (c) (~4.5 us).  For the 16-pixel-tile layers: a wide tile (less shared-memory traffic per output) on all SMs.
    if (g_csplit && !cell && !single && stacked && BN >= 64 && (halo_ok || taps == 1)) {
      const double tiles = (double)(halo_ok ? m_tiles_halo : m_tiles_tap) * tiles_n;
      const bool even = halo_ok ? (chunks % 2 == 0) : (chunks % 4 == 0);
      if (2 * tiles <= g_num_sms && even && ksteps_tile >= 32 && last_ksteps == kBK / 16) {
        const double bytes_tile = (halo_ok ? chunks * 2.0 * kHaloRows * 128 : items * 32768.0) + items * b_item;
        const double mma_tile = ksteps_tile * mpk * kMma;
        const double tile_t = mma_tile > bytes_tile / kSmBw ? mma_tile : bytes_tile / kSmBw;
        const double cost = kFixed + tile_t / 2 + 1200 + epi_tile;
        if (best.cost < 0 || cost < best.cost) {
          best = Plan{BN, stacked, 2, halo_ok ? 1 : 0, (long long)cost};
          best.csplit = 1;
        }
      }
    }
    // (b) split-K over (tap, chunk) items, TAP staging, single wave
    if (can_split && g_split_k) {
      const double tiles = (double)m_tiles_tap * tiles_n;
      int S = (int)(g_num_sms / tiles);
      if (S > items) S = items;
      if (S > 32) S = 32;
      if (tiles <= g_num_sms && S >= 2) {
        const double items_s = ceil_div(items, S);
        const double mma = items_s * (kBK / 16) * mpk * kMma;
        const double load = items_s * (32768.0 + b_item) / kSmBw;
        const double units_per_warp = ceil_div((kBM / 4) * ceil_div(BN, 32), kEpiWarps * S);
        double cost = kFixed + (mma > load ? mma : load) + 4500 + units_per_warp * (300 + 110.0 * S * (stacked ? 2 : 1));
        const double chip = kFixed + tiles * items * (32768.0 + b_item) / kChipBw;
        if (chip > cost) cost = chip;
        if (cost < best.cost) best = Plan{BN, stacked, S, 0, (long long)cost};
      }
    }
  }
  return best;
}

// Fills geometry, tensor maps and the K loop.
int setup(UmmaMaps& maps, UmmaParams& p, const rsis_tensor& x, const rsis_conv_weights* w, int stride, int pad,
          void* workspace, size_t workspace_bytes, int cta_share = 0, const float* preact = nullptr,
          bool is_cell = false) {
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
  p.chunks = ceil_div(x.c, kBK);
  p.last_ksteps = ceil_div(x.c - (p.chunks - 1) * kBK, 16);
  const bool halo_ok = g_halo_enabled && w->kh == 3 && stride == 1 && p.Wo % kHaloBW == 0 && p.Ho % kHaloBH == 0;
  // TAP-mode pixel box
  const int tBW = next_pow2(p.Wo) < kBM ? next_pow2(p.Wo) : kBM;
  const int tBH = next_pow2(p.Ho) < kBM / tBW ? next_pow2(p.Ho) : kBM / tBW;
  const int tBI = kBM / (tBW * tBH);
  const long long mt_tap = (long long)ceil_div(p.Wo, tBW) * ceil_div(p.Ho, tBH) * ceil_div(p.N, tBI);
  const long long mt_halo = (long long)(p.Wo / kHaloBW) * (p.Ho / kHaloBH) * p.N;
  if (mt_tap > 0x3fffffLL || mt_halo > 0x3fffffLL) return RSIS_ERR_UNSUPPORTED;
  const bool can_split = workspace != nullptr && workspace_bytes >= kWorkspaceBytes && aligned16(workspace);
  p.single = g_precision == 1 ? 1 : 0;
  p.early_b = (g_static_weights && g_early_b) ? 1 : 0;
#ifdef RSIS_DEBUG_TIMING
  p.dbg_skip = getenv("RSIS_B200_DBG_SKIP") ? atoi(getenv("RSIS_B200_DBG_SKIP")) : 0;
  p.trace = nullptr;
  if (g_trace && g_trace_next < g_trace_rows) p.trace = g_trace + (size_t)24 * g_trace_next++;
#endif
  const Plan plan = make_plan((int)mt_halo, (int)mt_tap, halo_ok, w->cout, p.taps, p.chunks, p.last_ksteps, can_split,
                              w->gate_interleaved != 0, cta_share, p.single != 0);
#ifdef RSIS_DEBUG_TIMING
  if (p.trace)
    fprintf(stderr, "rsis trace %d: N=%d %dx%d cin=%d cout=%d k=%d s=%d%s BN=%d ks=%d halo=%d tiles=%d\n", g_trace_next - 1, x.n,
            x.h, x.w, x.c, w->cout, w->kh, stride, is_cell ? " cell" : "", plan.BN, plan.ksplit, plan.halo,
            (int)((plan.halo ? mt_halo : mt_tap) * ceil_div(w->cout, plan.BN)) * plan.ksplit);
#endif
  if (g_print_plan)
    fprintf(stderr, "rsis plan: N=%d %dx%d cin=%d cout=%d k=%d s=%d%s -> BN=%d stacked=%d ksplit=%d%s halo=%d est %lld ns\n",
            x.n, x.h, x.w, x.c, w->cout, w->kh, stride, w->gate_interleaved ? " cell/gates" : "", plan.BN, plan.stacked,
            plan.ksplit, plan.csplit ? " (cluster)" : "", plan.halo, plan.cost);
  p.BN = plan.BN;
  p.stacked = plan.stacked;
  p.ksplit = plan.ksplit;
  p.halo = plan.halo;
  p.csplit = plan.csplit;
  p.x_off = 0;
  p.x_pitch = plan.BN + 4;
  if (p.halo) {
    p.BW = kHaloBW;
    p.BH = kHaloBH;
    p.BI = 1;
  } else {
    p.BW = tBW;
    p.BH = tBH;
    p.BI = tBI;
  }
  p.tiles_w = ceil_div(p.Wo, p.BW);
  p.tiles_h = ceil_div(p.Ho, p.BH);
  p.tiles_i = ceil_div(p.N, p.BI);
  p.tiles_n = ceil_div(w->cout, p.BN);
  const long long nt = (long long)p.tiles_w * p.tiles_h * p.tiles_i * p.tiles_n;
  if (nt * p.ksplit > 0x7fffffffLL) return RSIS_ERR_UNSUPPORTED;
  p.num_tiles = (int)nt;
  p.dn = make_fastdiv(p.tiles_n);
  p.dw = make_fastdiv(p.tiles_w);
  p.dh = make_fastdiv(p.tiles_h);
  p.dks = make_fastdiv(p.ksplit);
  if (can_split && getenv("RSIS_B200_DEBUG_TIMING")) g_debug_counters = reinterpret_cast<unsigned*>(workspace);
  if ((p.ksplit > 1 && !p.csplit) || (can_split && getenv("RSIS_B200_DEBUG_TIMING"))) {
    p.counters = reinterpret_cast<unsigned*>(workspace);
    p.scratch = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kCounterBytes);
  }
  const int a_rows = p.halo ? kHaloRows : kBM;
  p.a_sw64 = (g_sw64 && p.halo && x.c <= 32 && p.ksplit == 1) ? 1 : 0;
  const int planes = p.single ? 1 : 2;
  p.a_split_off = 0;
  {
    // split mode: only where it buys weight residency on the unrolled cell issue path (see UmmaParams::a_sw64)
    const int split_stage = round_up(2 * kHaloRows * (64 + 32), 1024);
    const long long w_bytes = (long long)p.taps * 2 * p.BN * 128;
    const bool many = p.num_tiles >= 2 * (cta_share > 0 ? cta_share : g_num_sms);
    if (g_sw64 && g_b_resident && is_cell && g_cell_rows && p.halo && p.ksplit == 1 && !p.single && p.stacked &&
        p.chunks == 1 && p.taps == 9 && x.c > 32 && x.c <= 48 && p.tiles_n == 1 && many &&
        2 * split_stage + w_bytes <= kDynSmem - 1023) {
      p.a_sw64 = 2;
      p.a_split_off = 2 * kHaloRows * 64;
    }
  }
  const int a_row_bytes = p.a_sw64 ? 64 : 128;
  p.a_plane_bytes = a_rows * a_row_bytes;
  p.a_chunk_bytes = p.a_sw64 == 2 ? round_up(2 * kHaloRows * (64 + 32), 1024) : round_up(planes * p.a_plane_bytes, 1024);
  p.a_sbo = p.halo ? (uint32_t)((kHaloBW + 2) * a_row_bytes) : 1024u;
  p.b_chunk_bytes = planes * p.BN * 128;
  // K blocks per TMA box (see UmmaParams::ag).  Split-K slices cut the walk anywhere, so they keep one block per box.
  p.ag = p.bg = 1;
  p.a_flat = 0;
  if (g_box_group && (p.ksplit == 1 || p.csplit)) {  // (the cluster split cuts the walk at a box boundary)
    if (p.halo) {
      p.bg = 3 * p.b_chunk_bytes <= 49152 ? 3 : 1;
    } else {
      // flat-pixel view: the pixel box is 128 consecutive pixels (full-width rows, whole images when it spans several)
      const bool flat = w->kh == 1 && stride == 1 && pad == 0 && x.c % kBK == 0 && tBW == p.Wo &&
                        (tBI == 1 || tBH == p.Ho) && (p.Ho % tBH == 0 || tBI > 1) && planes == 2;
      if (flat) {
        p.a_flat = 1;
        p.ag = p.chunks >= 2 ? 2 : 1;
      }
      p.bg = 32768 / p.b_chunk_bytes;
      if (p.bg < 1) p.bg = 1;
      if (p.bg > p.chunks) p.bg = p.chunks;
      if (p.bg > 4) p.bg = 4;
    }
  }
  p.pw = p.BN >= 64 ? 32 : 16;
  // pre_tma (cells with hoisted gates on the row-wise epilogue): two stages of BN/32 boxes of 16 KB take the place of
  // the transpose staging area, which that epilogue does not use
  p.pre_tma = 0;  // (retired with the per-unit distribution of the row-wise epilogue: it measured no gain, profiles/r2bm_*)
  p.p_stage_bytes = (p.BN >> 5) * 16384;
  // the row-wise cell epilogue needs no transpose staging area: its 33 KB go to the operand rings (a third activation
  // stage on the narrow levels, whose 46 KB halo boxes take ~2.5 us from issue to landing: profiles/r2l_group_stamps.txt)
  const bool rows_epi = is_cell && g_cell_rows && p.ksplit == 1;  // (a hoisted-gate CONVOLUTION also has gate-interleaved weights)
  const int x_bytes = 0;  // (the cluster split-K's exchange buffer, 128 x (BN + 4) floats, reuses the dead operand stages)
  const int budget = kDynSmem - 1023 - (p.pre_tma ? 2 * p.p_stage_bytes : (rows_epi ? 0 : kStageBytes)) - x_bytes;
  {
    // Weight residency: a persistent CTA that walks several pixel tiles of ONE output-channel tile re-reads the same
    // taps x chunks weight blocks for every tile; when they all fit next to two activation stages, load them once.
    const int items_b = p.taps * p.chunks;
    const bool many_tiles = p.num_tiles >= 2 * (cta_share > 0 ? cta_share : g_num_sms);
    if (g_b_resident && p.ksplit == 1 && p.tiles_n == 1 && many_tiles && items_b <= kMaxStages &&
        2 * p.ag * p.a_chunk_bytes + items_b * p.b_chunk_bytes <= budget) {
      p.b_resident = 1;
      // the whole set in as few boxes as the walk allows: HALO all nine taps of a chunk, TAP all chunks of a tap
      if (g_box_group) p.bg = p.halo ? p.taps : p.chunks;
    }
  }
  for (;;) {
    p.a_stage_bytes = p.ag * p.a_chunk_bytes;
    p.b_stage_bytes = p.bg * p.b_chunk_bytes;
    const int inner = p.halo ? p.taps : p.chunks, outer = p.halo ? p.chunks : p.taps;
    if (p.b_resident) {
      p.b_stages = outer * ceil_div(inner, p.bg);
      p.a_stages = (budget - p.b_stages * p.b_stage_bytes) / p.a_stage_bytes;
      if (p.a_stages > g_max_a_stages) p.a_stages = g_max_a_stages;
    } else if (p.halo) {
      // a third activation stage when enough weight stages still fit next to it: a halo box needs ~2.5 us from issue to
      // landing, longer than the MMAs of one tile on the narrow levels
      const int min_b = p.bg > 1 ? 2 : 4;
      p.a_stages = (budget - 3 * p.a_stage_bytes) / p.b_stage_bytes >= min_b ? 3 : 2;
      p.b_stages = (budget - p.a_stages * p.a_stage_bytes) / p.b_stage_bytes;
      if (p.b_stages > kMaxStages) {
        p.b_stages = kMaxStages;
        p.a_stages = (budget - p.b_stages * p.b_stage_bytes) / p.a_stage_bytes;
        if (p.a_stages > 3) p.a_stages = 3;
      }
    } else if (p.ag > 1 || p.bg > 1) {
      // grouped TAP boxes: two (64 KB) or three activation stages, the rest of the budget for weight stages
      p.a_stages = p.ag > 1 ? 2 : 3;
      p.b_stages = (budget - p.a_stages * p.a_stage_bytes) / p.b_stage_bytes;
      if (p.b_stages > kMaxStages) p.b_stages = kMaxStages;
    } else {
      p.a_stages = budget / (p.a_stage_bytes + p.b_stage_bytes);
      if (p.a_stages > kMaxStages) p.a_stages = kMaxStages;
      p.b_stages = p.a_stages;
    }
    if (p.a_stages >= 2 && p.b_stages >= 2) break;
    if (p.b_resident && p.a_stages >= 2) break;
    // does not fit: smaller boxes
    if (p.bg > 1 && !p.b_resident) p.bg = p.halo ? (p.bg == 9 ? 3 : 1) : p.bg / 2;
    else if (p.ag > 1) p.ag = 1;
    else if (p.b_resident) { p.b_resident = 0; p.bg = 1; }
    else break;
  }
  if (p.a_sw64 == 2 && !(p.b_resident && p.a_stages >= 2)) return RSIS_ERR_UNSUPPORTED;  // (chosen only where this holds)
  p.x_off = 0;  // start of the activation ring
  if (p.csplit && kBM * p.x_pitch * 4 > p.a_stages * p.a_stage_bytes + p.b_stages * p.b_stage_bytes) return RSIS_ERR_UNSUPPORTED;
  p.agroups = p.halo ? 1 : ceil_div(p.chunks, p.ag);
  p.a_tx_bytes = p.a_sw64 == 2 ? (uint32_t)(2 * kHaloRows * (64 + 32)) : (uint32_t)(p.ag * planes * p.a_plane_bytes);
  p.b_tx_bytes = (uint32_t)p.b_stage_bytes;
  if (p.a_stages < 1 || p.b_stages < 1) return RSIS_ERR_UNSUPPORTED;
  if (g_print_plan)
    fprintf(stderr, "rsis stages: ag=%d bg=%d flat=%d a_stages=%d x %d B, b_stages=%d x %d B, resident=%d\n", p.ag, p.bg,
            p.a_flat, p.a_stages, p.a_stage_bytes, p.b_stages, p.b_stage_bytes, p.b_resident);
  if (p.pre_tma) {
    // [N][Ho][Wo][Cout] float32 as {Cout, Wo, Ho, N}; box {32, BW, BH, BI} = the pixel tile, rows in TMEM lane order
    cuuint64_t dims[4] = {(cuuint64_t)p.Cout, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)p.N};
    cuuint64_t strides[3] = {(cuuint64_t)p.Cout * 4, (cuuint64_t)p.Wo * p.Cout * 4, (cuuint64_t)p.Ho * p.Wo * p.Cout * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)p.BW, (cuuint32_t)p.BH, (cuuint32_t)p.BI};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(&maps.pre, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(preact), dims, strides, box,
                          estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return RSIS_ERR_CUDA;
  }
  p.scale = w->scale;
  p.shift = w->shift;
  p.cell_rows = g_cell_rows;
  const int cout_pad = round_up(w->cout, 16);
  const int k_pad = p.taps * p.chunks * kBK;
  p.b_map3d = (p.bg == 1 && g_bmap3d) ? 1 : 0;
  if (p.b_map3d) {
    if (int e = encode_weight_map3(&maps.b, w->w_umma, cout_pad, k_pad, p.BN, planes)) return e;
  } else if (int e = encode_weight_map(&maps.b, w->w_umma, cout_pad, p.taps, p.chunks, p.BN, planes, p.halo ? p.bg : 1,
                                       p.halo ? 1 : p.bg)) {
    return e;
  }
  if (p.a_flat) {
    if (int e = encode_act_map_flat(&maps.a[0], x, p.ag, planes)) return e;
  } else if (stride == 1) {
    if (p.halo) {
      if (int e = encode_act_map(&maps.a[0], x, 1, 0, 0, kHaloBW + 2, kHaloBH + 2, 1, planes, p.a_sw64 ? 32 : kBK)) return e;
      if (p.a_sw64 == 2)  // channels 32..47 of the same tensor, 16 per row
        if (int e = encode_act_map(&maps.a[1], x, 1, 0, 0, kHaloBW + 2, kHaloBH + 2, 1, planes, 16)) return e;
    } else {
      if (int e = encode_act_map(&maps.a[0], x, 1, 0, 0, p.BW, p.BH, p.BI, planes)) return e;
    }
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        if (int e = encode_act_map(&maps.a[ph * 2 + pw], x, 2, ph, pw, p.BW, p.BH, p.BI, planes)) return e;
  }
  return RSIS_OK;
}

template <bool CELL, int PW, bool SPLIT, bool CS = false>
cudaError_t launch_one(const UmmaMaps& maps, const UmmaParams& p, int grid, cudaStream_t st, int cluster = 1) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreadsUmma);
  cfg.dynamicSmemBytes = kDynSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (g_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (cluster > 1) {  // cluster split-K: CTAs 2c and 2c + 1 are the two K slices of tile c
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = (unsigned)cluster;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, conv_umma_kernel<CELL, PW, SPLIT, CS>, maps, p);
}

thread_local int t_cta_cap = 0;  // > 0: at most this many CTAs for the next cell launch (set by convlstm_cell_umma)

template <bool CELL>
int launch(const UmmaMaps& maps, const UmmaParams& p, cudaStream_t st) {
  const int work = p.num_tiles * p.ksplit;
  int grid = work < g_num_sms ? work : g_num_sms;
  // a split-K launch needs all its slices co-resident; a plain persistent launch walks its tiles with any grid
  if (CELL && t_cta_cap > 0 && p.ksplit == 1 && grid > t_cta_cap) grid = t_cta_cap;
  cudaError_t e;
  if (p.csplit)  // one wave, every CTA one (tile, K half): the ordinary kernel with a cluster attribute
    e = p.pw == 32 ? launch_one<false, 32, false, true>(maps, p, work, st, 2)
                   : launch_one<false, 16, false, true>(maps, p, work, st, 2);
  else if (p.ksplit > 1)
    e = p.pw == 32 ? launch_one<CELL, 32, true>(maps, p, grid, st) : launch_one<CELL, 16, true>(maps, p, grid, st);
  else
    e = p.pw == 32 ? launch_one<CELL, 32, false>(maps, p, grid, st) : launch_one<CELL, 16, false>(maps, p, grid, st);
  RSIS_CUDA_TRY(e);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

// Weight tensor map for the swapped cell: the plane-interleaved pack [cout_pad][2 planes][k_pad] (rsis_conv_weights::
// w_umma_il) as {k, 2 * cout_pad rows}: one box {64, 2 * rows} lands as smem rows (gate column r, plane) -> 2r + plane.
int encode_weight_map_swapped(CUtensorMap* m, const void* w, int cout_pad, int k_pad, int rows) {
  cuuint64_t dims[3] = {(cuuint64_t)k_pad, (cuuint64_t)2 * cout_pad, 1};
  cuuint64_t strides[2] = {(cuuint64_t)k_pad * 2, (cuuint64_t)2 * cout_pad * k_pad * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)(2 * rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? RSIS_OK : RSIS_ERR_CUDA;
}

bool swap_eligible(const rsis_tensor& x, const rsis_conv_weights* w) {
  std::call_once(g_once, init_once);
  return g_swap && g_precision == 0 && g_init_status == RSIS_OK && w->w_umma_il && aligned16(w->w_umma_il) && w->kh == 3 &&
         (w->cout == 32 || w->cout == 64) && x.c <= kBK &&
         x.w % kHaloBW == 0 && x.h % kSwTileH == 0;
}

int setup_swap(UmmaMaps& maps, UmmaParams& p, const rsis_tensor& x, const rsis_conv_weights* w) {
  p.N = x.n;
  p.Ho = x.h;
  p.Wo = x.w;
  p.Cout = w->cout;
  p.BN = w->cout;
  p.BW = kHaloBW;
  p.BH = kSwTileH;
  p.BI = 1;
  p.tiles_w = x.w / kHaloBW;
  p.tiles_h = x.h / kSwTileH;
  p.tiles_i = x.n;
  p.tiles_n = 1;
  const long long nt = (long long)p.tiles_w * p.tiles_h * p.tiles_i;
  if (nt > 0x7fffffffLL) return RSIS_ERR_UNSUPPORTED;
  p.num_tiles = (int)nt;
  p.ksplit = 1;
  p.chunks = 1;
  p.taps = 9;
  p.last_ksteps = ceil_div(x.c, 16);
  p.b_stage_bytes = 2 * p.BN * 128;
  p.b_tx_bytes = (uint32_t)p.b_stage_bytes;
  // the M = 128 descriptor of a 64-row weight slot also reads the 64 rows behind it (ignored lanes): keep one slot
  // of slack behind the ring (the epilogue staging area follows, so the read stays inside this CTA's shared memory)
  // Shared memory: activation halo stages (85 KB each) + weight ring + epilogue staging (one slot per ACTIVE warp).
  // 32 gate columns: two activation stages (the next tile's halo loads under this tile's MMAs) and streamed weights;
  // 64 gate columns: one activation stage, as many weight stages as fit.
  const int stage_slots = 2 * p.BN >= 128 ? kEpiWarps : kEpiWarps / 2;
  p.a_stages = (p.BN == 32 && p.num_tiles > g_num_sms) ? 2 : 1;
  const int budget = kDynSmem - 1023 - stage_slots * kStageFloats * 4 - p.a_stages * kSwActBytes;
  p.b_stages = budget / p.b_stage_bytes;
  if (p.b_stages > kMaxStages) p.b_stages = kMaxStages;
  p.b_resident = (g_b_resident && p.b_stages >= 9 && p.num_tiles >= 2 * g_num_sms) ? 1 : 0;
  if (p.b_resident) p.b_stages = 9;
  if (p.b_stages < 2) return RSIS_ERR_UNSUPPORTED;
  p.scale = w->scale;
  p.shift = w->shift;
  const int cout_pad = round_up(w->cout, 16);
  const int k_pad = 9 * kBK;
  if (int e = encode_weight_map_swapped(&maps.b, w->w_umma_il, cout_pad, k_pad, p.BN)) return e;
  p.counters = g_debug_counters;
  if (int e = encode_act_map(&maps.a[0], x, 1, 0, 0, kHaloBW + 2, kSwTileH + 2, 1)) return e;
  return RSIS_OK;
}

int launch_swap(const UmmaMaps& maps, const UmmaParams& p, cudaStream_t st) {
  cudaLaunchConfig_t cfg{};
  int grid = p.num_tiles < g_num_sms ? p.num_tiles : g_num_sms;
  if (t_cta_cap > 0 && grid > t_cta_cap) grid = t_cta_cap;
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreadsUmma);
  cfg.dynamicSmemBytes = kDynSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  RSIS_CUDA_TRY(cudaLaunchKernelEx(&cfg, cell_swap_kernel, maps, p));
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}


// ===================================================================================================================
// tcgen05 weight gradient (the backward of every stride-1 1x1 / 3x3 convolution; train.py:184).
//   dW[tap][co][ci] = sum over output pixels p of dY[p][co] * X[p + tap offset][ci]
// GEMM view: D[M = 128 output channels][N = up to 4 x 64 (tap, input-channel chunk) columns] += A * B with the
// PIXELS as the contraction dimension.  NHWC activations have the channels contiguous, i.e. both operands are
// "MN-major": the smem tile TMA writes (rows = pixels, 64 bf16 channels = one 128-byte swizzle row) is consumed with
// the MN-major SWIZZLE_128B canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: 64 channels per chunk
// (chunks LBO bytes apart), K groups of 8 pixel rows (SBO = 1024 bytes apart); a_major = b_major = 1 in the
// instruction descriptor.  One 64-pixel box (TW x TH, TW*TH = 64) per stage and per operand chunk:
//   A  dY, split-bf16 planes: two boxes {64 co, TW, TH, 1, 2 planes} (co chunks 0/1 of the 128-row block);
//   B  X : up to four boxes, slot j = (tap, ci chunk) pair number 4*unit + j, the box origin shifted by the tap
//      offset (TMA zero-fills outside the map = the convolution's zero padding, and beyond the last channel).
// Three kind::f16 MMAs (128 x N x 16) per 16 pixels: dY_hi*X_hi + dY_hi*X_lo + dY_lo*X_hi, fp32 accumulation in TMEM.
// Grid: (N unit, co block, pixel split); the partial dW tiles of the pixel splits are combined with 16-byte vector
// reductions (red.global.add.v4.f32) into a zero-initialised scratch [tap][co][ci_pad], which a finishing kernel
// transposes into the reference's OIHW layout (and re-zeroes).
// Warps: 0 TMA producer, 1 MMA issuer (+ TMEM allocation), 2-5 epilogue (TMEM lane quarter = warp % 4).
// ===================================================================================================================
constexpr int kWgTilePx = 64;
constexpr int kWgPlaneBytes = kWgTilePx * 128;      // one bf16 plane of one 64-channel chunk: 8 KB
constexpr int kWgChunkBytes = 2 * kWgPlaneBytes;    // hi|lo planes = one TMA box: 16 KB
constexpr int kWgAStage = 2 * kWgChunkBytes;        // two co chunks
constexpr int kWgBStage = 4 * kWgChunkBytes;        // four (tap, ci chunk) slots
constexpr int kWgStageBytes = kWgAStage + kWgBStage;  // 96 KB
constexpr int kWgStages = 2;
constexpr int kWgThreads = 192;
constexpr int kWgTmemCols = 256;
constexpr int kWgDynSmem = kWgStages * kWgStageBytes + 1024;

struct alignas(64) WgMaps {
  CUtensorMap dy;
  CUtensorMap x[4];  // stride 1: x[0]; stride 2: one per (row parity, column parity) sub-grid of the input
};

struct WgParams {
  int TW, TH;
  int tiles_w, tiles_h, num_tiles;
  int tiles_per_cta, splits;
  int units_n, co_blocks;
  int pairs, chunks, ksize, pad, stride;
  int Cout, ci_pad;
  int single;  // single-pass bf16: dY_hi * X_hi only
  float* scratch;
};

// MN-major SWIZZLE_128B shared-memory matrix descriptor (see the block comment above).
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;   // bytes between 64-element chunks along M / N
  d |= (uint64_t)(sbo >> 4) << 32;   // bytes between groups of 8 K rows (pixels)
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;            // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_umma_kernel(const __grid_constant__ WgMaps maps, const WgParams p) {
  extern __shared__ __align__(1024) uint8_t wg_smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * kWgStages + 1];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = (smem_u32(wg_smem_raw) + 1023u) & ~1023u;
  const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[kWgStages]), done = smem_u32(&bars[2 * kWgStages]);

  // CTA -> (pixel split, N unit, co block)
  const int s = blockIdx.x % p.splits;
  const int u = blockIdx.x / p.splits;
  const int nu = u % p.units_n;
  const int cb = u / p.units_n;
  const int co0 = cb * 128;
  const int n_co_chunks = (p.Cout - co0) > 64 ? 2 : 1;
  const int pair0 = nu * 4;
  const int n_slots = (p.pairs - pair0) < 4 ? (p.pairs - pair0) : 4;
  const int t_begin = s * p.tiles_per_cta;
  const int t_end = (t_begin + p.tiles_per_cta) < p.num_tiles ? (t_begin + p.tiles_per_cta) : p.num_tiles;
  const int my_tiles = t_end - t_begin;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.dy);
    prefetch_tmap(&maps.x[0]);
    if (p.stride == 2) {
      prefetch_tmap(&maps.x[1]);
      prefetch_tmap(&maps.x[2]);
      prefetch_tmap(&maps.x[3]);
    }
    for (int i = 0; i < kWgStages; ++i) {
      mbar_init(full0 + 8 * i, 1);
      mbar_init(empty0 + 8 * i, 1);
    }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "n"(kWgTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_trigger();
  pdl_wait();

  if (my_tiles > 0) {
    if (warp == 0) {
      // =============================== TMA producer ===============================
      if (elect_one()) {
        const uint32_t tx = (uint32_t)(n_co_chunks + n_slots) * kWgChunkBytes;
        const int tiles_img = p.tiles_w * p.tiles_h;
        for (int it = 0; it < my_tiles; ++it) {
          const int st = it % kWgStages;
          mbar_wait(empty0 + 8 * st, ((it / kWgStages) & 1) ^ 1);
          const int t = t_begin + it;
          const int n = t / tiles_img;
          const int r = t - n * tiles_img;
          const int ty = r / p.tiles_w, txi = r - ty * p.tiles_w;
          const int y0 = ty * p.TH, x0 = txi * p.TW;
          const uint32_t sa = smem0 + st * kWgStageBytes, sb = sa + kWgAStage;
          mbar_arrive_expect_tx(full0 + 8 * st, tx);
          for (int cc = 0; cc < n_co_chunks; ++cc)
            tma_load_5d(sa + cc * kWgChunkBytes, &maps.dy, full0 + 8 * st, co0 + cc * 64, x0, y0, n, 0);
          for (int j = 0; j < n_slots; ++j) {
            const int pair = pair0 + j;
            const int tap = pair / p.chunks, chunk = pair - tap * p.chunks;
            const int kh = tap / p.ksize, kw = tap - kh * p.ksize;
            // input pixel of output pixel (y, x) under this tap: (stride*y + kh - pad, stride*x + kw - pad).  Stride 2:
            // that is pixel (y + sy, x + sx) of the (row parity, column parity) sub-grid map.
            int oy = kh - p.pad, ox = kw - p.pad, mi = 0;
            if (p.stride == 2) {
              const int ph = oy & 1, pw = ox & 1;
              oy = (oy - ph) >> 1;
              ox = (ox - pw) >> 1;
              mi = ph * 2 + pw;
            }
            tma_load_5d(sb + j * kWgChunkBytes, &maps.x[mi], full0 + 8 * st, chunk * 64, x0 + ox, y0 + oy, n, 0);
          }
        }
      }
    } else if (warp == 1) {
      // =============================== MMA issuer ===============================
      if (elect_one()) {
        const uint32_t n_mma = (uint32_t)n_slots * 64u;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((n_mma >> 3) << 17) |
                               ((128u >> 4) << 24);
        for (int it = 0; it < my_tiles; ++it) {
          const int st = it % kWgStages;
          mbar_wait(full0 + 8 * st, (it / kWgStages) & 1);
          tc_fence_after();
          const uint32_t sa = smem0 + st * kWgStageBytes, sb = sa + kWgAStage;
          const uint64_t a_hi0 = make_smem_desc_mn(sa, kWgChunkBytes, 1024);
          const uint64_t b_hi0 = make_smem_desc_mn(sb, kWgChunkBytes, 1024);
#pragma unroll
          for (int ks = 0; ks < kWgTilePx / 16; ++ks) {
            const uint64_t adv = (uint64_t)((ks * 16 * 128) >> 4);  // 16 pixel rows of 128 bytes
            const uint64_t a_hi = a_hi0 + adv, a_lo = a_hi + (kWgPlaneBytes >> 4);
            const uint64_t b_hi = b_hi0 + adv, b_lo = b_hi + (kWgPlaneBytes >> 4);
            umma_bf16(tmem_base, a_hi, b_hi, idesc, (it | ks) ? 1u : 0u);
            if (!p.single) {
              umma_bf16(tmem_base, a_hi, b_lo, idesc, 1u);
              umma_bf16(tmem_base, a_lo, b_hi, idesc, 1u);
            }
          }
          umma_commit(empty0 + 8 * st);  // the stage is free once these MMAs have read it
        }
        umma_commit(done);
      }
    } else {
      // =============================== epilogue (warps 2-5) ===============================
      const int q = warp & 3;
      mbar_wait(done, 0);
      tc_fence_after();
      const int co = co0 + 32 * q + lane;
      const bool row_ok = co < p.Cout;
      if (32 * q < (n_co_chunks * 64)) {
        for (int j = 0; j < n_slots; ++j) {
          const int pair = pair0 + j;
          const int tap = pair / p.chunks, chunk = pair - tap * p.chunks;
          float* dst = p.scratch + ((size_t)tap * p.Cout + (row_ok ? co : 0)) * p.ci_pad + chunk * 64;
#pragma unroll
          for (int piece = 0; piece < 2; ++piece) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(j * 64 + piece * 32), r);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int v = 0; v < 8; ++v)
                red_add_v4(dst + piece * 32 + 4 * v, __uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                           __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kWgTmemCols) : "memory");
  }
}

// scratch [taps][Cout][ci_pad] -> dw OIHW [Cout][Cin][taps] (overwrite or accumulate); leaves the scratch zeroed.
// One block per (output channel, slice of the 64-channel chunks): a chunk's taps x 64 values are read as `taps`
// contiguous 256-byte rows, transposed through shared memory and written as one contiguous run of the OIHW tensor.
__global__ void __launch_bounds__(256) wgrad_finish_kernel(float* __restrict__ scratch, float* __restrict__ dw, int taps,
                                                           int Cout, int Cin, int ci_pad, int accumulate) {
  __shared__ float tile[9 * 64];
  const int co = blockIdx.x;
  const int chunks = ci_pad / 64;
  for (int ch = blockIdx.y; ch < chunks; ch += gridDim.y) {
    const int c0 = ch * 64;
    for (int i = threadIdx.x; i < taps * 64; i += blockDim.x) {
      const int tap = i >> 6, j = i & 63;
      float* src = scratch + ((size_t)tap * Cout + co) * ci_pad + c0 + j;
      tile[j * taps + tap] = *src;
      *src = 0.f;
    }
    __syncthreads();
    const int valid = (Cin - c0 < 64 ? Cin - c0 : 64) * taps;
    float* dst = dw + ((size_t)co * Cin + c0) * taps;
    for (int i = threadIdx.x; i < valid; i += blockDim.x) dst[i] = accumulate ? dst[i] + tile[i] : tile[i];
    __syncthreads();
  }
}

}  // namespace

size_t conv_wgrad_umma_workspace_bytes() { return (size_t)4 * 1024 * 1024 * sizeof(float); }

bool conv_wgrad_umma_supported(const rsis_tensor* x, const rsis_tensor* dy, int kh, int kw, int stride, int pad,
                               size_t workspace_bytes) {
  std::call_once(g_once, init_once);
  if (g_init_status != RSIS_OK || !x || !dy) return false;
  if (kh != kw || (kh != 1 && kh != 3) || (stride != 1 && stride != 2) || pad != kh / 2) return false;
  if (!split_ok(x) || !split_ok(dy) || dy->c % 8 != 0) return false;
  if (stride == 2 && ((x->h & 1) || (x->w & 1))) return false;
  if (x->n != dy->n || x->h != dy->h * stride || x->w != dy->w * stride) return false;
  const size_t need = (size_t)kh * kw * dy->c * round_up(x->c, 64) * sizeof(float);
  return need <= workspace_bytes && need <= conv_wgrad_umma_workspace_bytes();
}

// dw_oihw (+)= wgrad(x, dy).  workspace: conv_wgrad_umma_workspace_bytes() bytes, zero-filled once by the caller (the
// finishing kernel re-zeroes what it used).
int conv_wgrad_umma(const rsis_tensor* x, const rsis_tensor* dy, int ksize, int stride, float* dw_oihw, int accumulate,
                    void* workspace, cudaStream_t st) {
  static std::once_flag attr_once;
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(attr_once, [] {
    attr_err = cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgDynSmem);
  });
  if (attr_err != cudaSuccess) {
    set_cuda_error(attr_err);
    return RSIS_ERR_CUDA;
  }
  WgMaps maps;
  WgParams p{};
  p.TW = dy->w > 8 ? 16 : 8;   // tiles walk the OUTPUT pixels
  p.TH = kWgTilePx / p.TW;
  p.tiles_w = ceil_div(dy->w, p.TW);
  p.tiles_h = ceil_div(dy->h, p.TH);
  p.stride = stride;
  const long long nt = (long long)dy->n * p.tiles_w * p.tiles_h;
  if (nt > 0x7fffffffLL) return RSIS_ERR_UNSUPPORTED;
  p.num_tiles = (int)nt;
  p.chunks = ceil_div(x->c, 64);
  p.ksize = ksize;
  p.pad = ksize / 2;
  p.pairs = ksize * ksize * p.chunks;
  p.units_n = ceil_div(p.pairs, 4);
  p.co_blocks = ceil_div(dy->c, 128);
  p.Cout = dy->c;
  p.ci_pad = p.chunks * 64;
  p.scratch = reinterpret_cast<float*>(workspace);
  p.single = g_precision == 1 ? 1 : 0;
  // 1x1 convolutions whose Cin is a multiple of 64: the scratch layout [co][ci] IS the OIHW tensor -- reduce straight
  // into the gradient (no finishing pass)
  const bool direct = ksize == 1 && x->c % 64 == 0 && aligned16(dw_oihw);
  if (direct) {
    p.scratch = dw_oihw;
    if (!accumulate) RSIS_CUDA_TRY(cudaMemsetAsync(dw_oihw, 0, (size_t)p.Cout * x->c * sizeof(float), st));
  }
  const long long units = (long long)p.units_n * p.co_blocks;
  // one CTA per SM (192 KB of operand stages): split the pixels so that the launch is ONE wave
  long long splits = units >= g_num_sms ? 1 : g_num_sms / units;
  if (splits > p.num_tiles) splits = p.num_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_cta = (int)((p.num_tiles + splits - 1) / splits);
  p.splits = ceil_div(p.num_tiles, p.tiles_per_cta);
  if (units * p.splits > 0x7fffffffLL) return RSIS_ERR_UNSUPPORTED;
  if (int e = encode_act_map(&maps.dy, *dy, 1, 0, 0, p.TW, p.TH, 1)) return e;
  if (stride == 1) {
    if (int e = encode_act_map(&maps.x[0], *x, 1, 0, 0, p.TW, p.TH, 1)) return e;
  } else {
    for (int ph = 0; ph < 2; ++ph)
      for (int pw = 0; pw < 2; ++pw)
        if (int e = encode_act_map(&maps.x[ph * 2 + pw], *x, 2, ph, pw, p.TW, p.TH, 1)) return e;
  }
  wgrad_umma_kernel<<<(unsigned)(units * p.splits), kWgThreads, kWgDynSmem, st>>>(maps, p);
  RSIS_CHECK_LAUNCH();
  if (direct) return RSIS_OK;
  // one block per (output channel, 64-channel chunk): a single load -> transpose -> store round trip per block
  int ysplit = p.chunks < 65535 ? p.chunks : 65535;
  wgrad_finish_kernel<<<dim3((unsigned)p.Cout, (unsigned)ysplit), 256, 0, st>>>(p.scratch, dw_oihw, ksize * ksize, p.Cout,
                                                                              x->c, p.ci_pad, accumulate);
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

bool conv2d_umma_supported(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const rsis_tensor* residual,
                           const rsis_tensor* y, const rsis_tensor* y2, int stride, int pad) {
  if (n_src != 1 || !common_supported(srcs, w, stride, pad)) return false;
  if (!out_ok(y)) return false;
  if (y2 && !out_ok(y2)) return false;
  if (residual && !out_ok(residual)) return false;
  return true;
}

size_t conv_umma_workspace_bytes() { return kWorkspaceBytes; }

int conv2d_umma(const rsis_tensor* srcs, int n_src, const rsis_conv_weights* w, const rsis_tensor* residual,
                const rsis_tensor* y, const rsis_tensor* y2, int stride, int pad, int relu, void* workspace,
                size_t workspace_bytes, cudaStream_t st) {
  (void)n_src;
  UmmaMaps maps;
  UmmaParams p{};
  if (int e = setup(maps, p, srcs[0], w, stride, pad, workspace, workspace_bytes)) return e;
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
                       const float* gate_preact, const rsis_tensor* h_out, const rsis_tensor* h_split, const rsis_tensor* c_out,
                       uint32_t* side_max, int side_stride, int side_offset, void* workspace, size_t workspace_bytes,
                       int cta_cap, cudaStream_t st) {
  (void)n_src;
  struct CapGuard {
    explicit CapGuard(int c) { t_cta_cap = c; }
    ~CapGuard() { t_cta_cap = 0; }
  } cap_guard(cta_cap);
  UmmaMaps maps;
  UmmaParams p{};
  const bool swapped = swap_eligible(srcs[0], w);
  if (swapped) {
    if (int e = setup_swap(maps, p, srcs[0], w)) return e;
  } else {
    if (gate_preact && !aligned16(gate_preact)) return RSIS_ERR_ALIGN;
    if (int e = setup(maps, p, srcs[0], w, 1, w->kh / 2, workspace, workspace_bytes, 0, gate_preact, true)) return e;
  }
  const int Ch = p.Cout / 4;
  auto ok = [&](const rsis_tensor* t, int fmt) {
    return valid_tensor(t) && t->fmt == fmt && t->n == p.N && t->h == p.Ho && t->w == p.Wo && t->c == Ch &&
           aligned16(t->data);
  };
  if (!ok(h_out, RSIS_FMT_F32) || !ok(c_out, RSIS_FMT_F32) || pitch(*h_out) != Ch || pitch(*c_out) != Ch)
    return RSIS_ERR_BAD_ARG;
  if (h_split && (!ok(h_split, RSIS_FMT_SPLIT_BF16) || pitch(*h_split) % 8 != 0)) return RSIS_ERR_BAD_ARG;
  if (side_max && (side_stride < side_offset + Ch || side_offset < 0)) return RSIS_ERR_BAD_ARG;
  if ((c_prev && !aligned16(c_prev)) || (gate_preact && !aligned16(gate_preact))) return RSIS_ERR_ALIGN;
  p.c_prev = c_prev;
  p.preact = gate_preact;
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
  if (swapped) return launch_swap(maps, p, st);
  return launch<true>(maps, p, st);
}


// ---- precision mode -----------------------------------------------------------------------------------------------
int set_precision(int mode) {
  const int prev = g_precision;
  if (mode == 0 || mode == 1) g_precision = mode;
  return prev;
}
int get_precision() { return g_precision; }
int set_static_weights(int on) {
  const int prev = g_static_weights;
  g_static_weights = on != 0;
  return prev;
}

// ---- grouped cells: the cells of one wavefront of the decoder in one launch ----------------------------------------
int convlstm_cell_group_max() { return kMaxGroup; }

bool convlstm_cell_group_supported(const rsis_cell_args* cells, int n) {
  if (!cells || n < 1 || n > kMaxGroup) return false;
  for (int i = 0; i < n; ++i) {
    const rsis_cell_args& c = cells[i];
    if (!c.x || !c.w || !c.h_out || !c.c_out || !convlstm_cell_umma_supported(c.x, 1, c.w)) return false;
  }
  return true;
}

int convlstm_cell_group_umma(const rsis_cell_args* cells, int n, cudaStream_t st) {
  std::call_once(g_once, init_once);
  if (g_init_status != RSIS_OK) return g_init_status;
  if (!g_cell_rows) return RSIS_ERR_UNSUPPORTED;  // the grouped kernel carries the row-wise epilogue only
  static_assert(sizeof(CellGroup) < 32000, "kernel parameter space");
  CellGroup g{};
  g.n = n;
  // (1) CTA shares.  Model of a cell's time on n CTAs, calibrated with the cold single-launch sweep of
  // scripts/group_tune.py (profiles/r2bl_group_tune.txt): F + W / n, W proportional to its tensor-core work (pixel tiles x
  // K steps x gate columns, with the shared-memory-bound cost of narrow N), F an extra fixed cost of the levels whose
  // weights stream through the B ring for every tile (too large to stay resident, too few gate columns to split).
  // Greedy min-max: every cell one CTA, each further CTA to the cell that currently finishes last.
  double work[kMaxGroup], fixed[kMaxGroup];
  int tiles[kMaxGroup], share[kMaxGroup];
  for (int i = 0; i < n; ++i) {
    const rsis_tensor& x = *cells[i].x;
    const rsis_conv_weights* w = cells[i].w;
    tiles[i] = ceil_div(x.n * x.h * x.w, kBM);
    const double cols = w->cout < 16 ? 16 : w->cout;
    // narrow gate blocks are bound by the shared-memory operand reads, not by the tensor pipe: 32 + N/4 vs N/2 cycles
    // (re-fitted after the three-warps-per-quarter epilogue and the narrow activation boxes: level 4 on 61 CTAs 42 us,
    // level 3 on 35 CTAs 40 us, profiles/r2by_group_tune.txt)
    const double per_col = cols >= 128 ? (cols == 256 ? 0.67 : 1.0) : (35.8 + 0.29 * cols) / cols;  // (fitted: r2bl, r2dd)
    work[i] = (double)tiles[i] * w->kh * w->kw * ceil_div(x.c, 16) * cols * per_col * 2.06e-3;  // ~us on one CTA
    const double w_bytes = 2.0 * w->kh * w->kw * ceil_div(x.c, kBK) * 128.0 * cols;              // [W_hi | W_lo] of a tile
    const double a_stage = x.c <= 32 ? 23552.0 : (x.c <= 48 ? 34816.0 : 46080.0 * ceil_div(x.c, kBK));
    const bool can_reside = x.c <= kBK && 2 * a_stage + w_bytes <= kDynSmem - 1023;
    fixed[i] = (cols < 128 && !can_reside && tiles[i] >= 2 * g_num_sms / n) ? 5.5 : 0.0;
    tiles[i] *= ceil_div(w->cout, 32);  // most (pixel tile, gate-column tile) units a plan can have
    share[i] = 1;
  }
  for (int used = n; used < g_num_sms; ++used) {
    int best = -1;
    double worst = -1;
    for (int i = 0; i < n; ++i) {
      if (share[i] >= tiles[i]) continue;
      const double t = fixed[i] + work[i] / share[i];
      if (t > worst) { worst = t; best = i; }
    }
    if (best < 0) break;
    ++share[best];
  }
  // (1b) quantisation: a cell with few work units (pixel tile x gate-column tile of the plan its share gets) runs
  // ceil(units / share) rounds whatever the remainder, so it keeps only the CTAs that round count needs and the
  // many-tile cells of the group (if there are any) take the rest.  One planning pass to learn the unit counts.
  {
    int need[kMaxGroup], freed = 0, n_big = 0;
    bool big[kMaxGroup];
    for (int i = 0; i < n; ++i) {
      UmmaMaps maps_tmp;
      UmmaParams p_tmp{};
      const rsis_cell_args& c = cells[i];
      big[i] = true;
      need[i] = share[i];
      if (setup(maps_tmp, p_tmp, *c.x, c.w, 1, c.w->kh / 2, nullptr, 0, share[i], c.gate_preact, true) == RSIS_OK) {
        const int units = p_tmp.num_tiles;
        big[i] = units >= 4 * share[i];
        if (!big[i]) need[i] = ceil_div(units, ceil_div(units, share[i]));
      }
      if (big[i]) ++n_big;
    }
    if (n_big > 0) {
      for (int i = 0; i < n; ++i) {
        freed += share[i] - need[i];
        share[i] = need[i];
      }
      int total = 0;
      for (int i = 0; i < n; ++i) total += share[i];
      freed = g_num_sms - total;  // (negative when (1a) rounded shares up)
      for (; freed > 0; --freed) {
        int best = -1;
        double worst = -1;
        for (int i = 0; i < n; ++i) {
          if (!big[i] || share[i] >= tiles[i]) continue;
          const double t = fixed[i] + work[i] / share[i];
          if (t > worst) { worst = t; best = i; }
        }
        if (best < 0) break;
        ++share[best];
      }
      for (; freed < 0; ++freed) {  // take back from the many-tile cell that finishes first
        int best = -1;
        double least = 1e300;
        for (int i = 0; i < n; ++i) {
          if (!big[i] || share[i] <= 1) continue;
          const double t = fixed[i] + work[i] / share[i];
          if (t < least) { least = t; best = i; }
        }
        if (best < 0) break;
        --share[best];
      }
    }
  }
  // (1c') a small cell's plan changes with its share: with 4 pixel tiles and 512 gate columns, 8 CTAs get 8 units of 256
  // columns (a 21 us MMA chain and a 10 us epilogue each) where 16 CTAs get 16 units of 128 -- 46 -> 32 us for that cell
  // alone, and 49 -> 35 us for the wavefront of levels 0-3 it was holding up (profiles/r2dd_group_tune_partial.txt).  When
  // the cell that finishes last by the model (by a margin) is such a cell, it moves up to its next unit count and the
  // cells that finish first pay for it.
  for (int round = 0; round < 2; ++round) {
    int m = -1, second = -1;
    double tm = -1, ts = -1;
    for (int i = 0; i < n; ++i) {
      const double t = fixed[i] + work[i] / share[i];
      if (t > tm) { ts = tm; second = m; tm = t; m = i; }
      else if (t > ts) { ts = t; second = i; }
    }
    (void)second;
    if (m < 0 || n < 2 || tm < 1.25 * ts) break;
    const rsis_tensor& x = *cells[m].x;
    const int tiles_px = ceil_div(x.n * x.h * x.w, kBM), cout = cells[m].w->cout;
    int cand = 0;
    for (int b = kMaxBN; b >= 32; b >>= 1) {
      const int c = tiles_px * ceil_div(cout, b);
      if (c > share[m] && c <= 2 * share[m] + 1 && c <= g_num_sms / 2) { cand = c; break; }
    }
    if (!cand) break;
    int saved[kMaxGroup];
    for (int i = 0; i < n; ++i) saved[i] = share[i];
    int need_ctas = cand - share[m];
    share[m] = cand;
    while (need_ctas > 0) {  // from the cell that finishes first
      int best = -1;
      double least = 1e300;
      for (int i = 0; i < n; ++i) {
        if (i == m || share[i] <= 1) continue;
        const double t = fixed[i] + work[i] / (share[i] - 1);
        if (t < least) { least = t; best = i; }
      }
      if (best < 0) break;
      --share[best];
      --need_ctas;
    }
    double new_max = 0;
    for (int i = 0; i < n; ++i) new_max = fmax(new_max, fixed[i] + work[i] / share[i]);
    if (new_max > 0.85 * tm) {  // the others would pay as much as this cell gains: keep the shares
      for (int i = 0; i < n; ++i) share[i] = saved[i];
      break;
    }
  }
  // (1c) never more CTAs than SMs (one wave): whatever was added beyond that comes back from the largest shares
  for (;;) {
    int total = 0, big_i = 0;
    for (int i = 0; i < n; ++i) {
      total += share[i];
      if (share[i] > share[big_i]) big_i = i;
    }
    if (total <= g_num_sms || share[big_i] <= 1) break;
    --share[big_i];
  }
  // tuning overrides (development): RSIS_B200_GROUP_SHARES / RSIS_B200_GROUP_BN = comma lists indexed by the cell's
  // position in the group
  int force_bn[kMaxGroup] = {0, 0, 0, 0, 0};
  if (const char* e = getenv("RSIS_B200_GROUP_SHARES")) {
    int v[kMaxGroup], k = 0;
    for (const char* q = e; *q && k < kMaxGroup; ++k) {
      v[k] = atoi(q);
      while (*q && *q != ',') ++q;
      if (*q == ',') ++q;
    }
    if (k == n) for (int i = 0; i < n; ++i) share[i] = v[i] < 1 ? 1 : v[i];
  }
  if (const char* e = getenv("RSIS_B200_GROUP_BN")) {
    int k = 0;
    for (const char* q = e; *q && k < kMaxGroup; ++k) {
      force_bn[k] = atoi(q);
      while (*q && *q != ',') ++q;
      if (*q == ',') ++q;
    }
    if (k != n) for (int i = 0; i < kMaxGroup; ++i) force_bn[i] = 0;
  }
  // (2) plan every cell for its share; no split-K (the wavefront supplies the parallelism)
  int first = 0;
  for (int i = 0; i < n; ++i) {
    const rsis_cell_args& c = cells[i];
    UmmaParams& p = g.p[i];
    if (c.gate_preact && !aligned16(c.gate_preact)) return RSIS_ERR_ALIGN;
    t_force_bn = force_bn[i];
    const int se = setup(g.maps[i], p, *c.x, c.w, 1, c.w->kh / 2, nullptr, 0, share[i], c.gate_preact, true);
    t_force_bn = 0;
    if (se) return se;
    const int Ch = p.Cout / 4;
    auto ok = [&](const rsis_tensor* t, int fmt) {
      return valid_tensor(t) && t->fmt == fmt && t->n == p.N && t->h == p.Ho && t->w == p.Wo && t->c == Ch &&
             aligned16(t->data);
    };
    if (!ok(c.h_out, RSIS_FMT_F32) || !ok(c.c_out, RSIS_FMT_F32) || pitch(*c.h_out) != Ch || pitch(*c.c_out) != Ch)
      return RSIS_ERR_BAD_ARG;
    if (c.h_split && (!ok(c.h_split, RSIS_FMT_SPLIT_BF16) || pitch(*c.h_split) % 8 != 0)) return RSIS_ERR_BAD_ARG;
    if (c.side_max && (c.side_stride < c.side_offset + Ch || c.side_offset < 0)) return RSIS_ERR_BAD_ARG;
    if ((c.c_prev && !aligned16(c.c_prev)) || (c.gate_preact && !aligned16(c.gate_preact))) return RSIS_ERR_ALIGN;
    p.c_prev = c.c_prev;
    p.preact = c.gate_preact;
    p.h_out = reinterpret_cast<float*>(c.h_out->data);
    p.c_out = reinterpret_cast<float*>(c.c_out->data);
    if (c.h_split) {
      p.h_split = reinterpret_cast<__nv_bfloat16*>(c.h_split->data);
      p.hs_cs = pitch(*c.h_split);
      p.hs_plane = plane_elems(*c.h_split);
    }
    p.side_max = c.side_max;
    p.side_stride = c.side_stride;
    p.side_offset = c.side_offset;
    if (getenv("RSIS_B200_DEBUG_TIMING")) p.counters = g_debug_counters;  // in-kernel stamps of block 0 (debug builds)
    int ctas = share[i] < p.num_tiles ? share[i] : p.num_tiles;
    if (ctas < 1) ctas = 1;
    g.first[i] = first;
    first += ctas;
    if (g_print_plan) fprintf(stderr, "rsis group: cell %d -> %d CTAs for %d tiles\n", i, ctas, p.num_tiles);
  }
  for (int i = n; i <= kMaxGroup; ++i) g.first[i] = first;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(first);
  cfg.blockDim = dim3(kThreadsGroup);
  cfg.dynamicSmemBytes = kDynSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_pdl ? 1 : 0;
  RSIS_CUDA_TRY(cudaLaunchKernelEx(&cfg, cell_group_kernel, g));
  RSIS_CHECK_LAUNCH();
  return RSIS_OK;
}

}  // namespace rsis

#ifdef RSIS_DEBUG_TIMING
// Debug builds only (not part of the ABI): every tcgen05 launch set up from now on gets a trace row in `buf`
// (rows x 24 x u64, zeroed by the caller): [0] ~(earliest CTA start), [1] latest CTA end (%globaltimer, ns),
// [2..17] block 0's stamps 0..15 (SM cycles), [18] / [19] block 0's start as %globaltimer / cycles, [20] / [21] its end.
extern "C" void rsis_debug_trace(void* buf, int rows) {
  rsis::g_trace = reinterpret_cast<unsigned long long*>(buf);
  rsis::g_trace_rows = rows;
  rsis::g_trace_next = 0;
}
#endif
