// espresso_b200 -- persistent warp-specialised bf16 GEMM for sm_100a.
//
//   C[b][M,N] = epilogue( op(A[b])[M,K] * op(B[b])[N,K]^T )        fp32 accumulate in TMEM
//
// * operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_128B) into a multi-stage smem ring,
// * tcgen05.mma (cta_group::1, kind::f16, 128 x BN x 16) issued by one elected thread,
// * accumulators double-buffered in TMEM so the epilogue of tile i overlaps the mainloop of i+1,
// * epilogue warps read TMEM with tcgen05.ld and fuse bias / activation (+grad) / dropout /
//   scaled residual / relative-position skew, writing bf16 or fp32 straight to HBM.
//
// This one kernel replaces every dense contraction on the Espresso hot path:
//   Linear layers                fairseq/modules/conformer_layer.py:134-146 (FFN), :79-101 (pw convs)
//   q/k/v/out/pos projections    fairseq/modules/multihead_attention.py:650-653,799-814,899-907
//   QK^T, (q+v)P^T, PV           fairseq/modules/multihead_attention.py:788,815-823,897
//   fc0 / fc_out                 espresso/models/transformer/speech_transformer_encoder.py:341-343,
//                                .../speech_transformer_encoder_model.py:207-208
// and their dgrad / wgrad counterparts (operand majors are template flags: K-major or MN-major
// smem descriptors, so no transposes are ever materialised).
#include "common.cuh"
#include "espresso_b200.h"
#include <cuda.h>
#include <stdlib.h>

void esp_count_launch(int n);

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int kThreads = 384;  // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-11 epilogue (2 per TMEM lane quarter)
constexpr int kEpiWarps = 8;
constexpr int kAccStages = 2;

struct EpiParams {
  void* C;
  bf16* C2;
  const bf16* bias;
  const bf16* aux;
  const void* R;
  long ldc, ld_aux, ldr;
  long sC1, sC2, sAux1, sAux2, sR1, sR2;
  int c_f32, r_f32, act, drop_mode, skew_r, atomic;
  float alpha, beta, drop_scale;
  uint32_t drop_thresh;
  unsigned long long seed;
  const unsigned long long* seed_ptr;
};

// 3x3 convolutions of the conv front end as implicit GEMMs on this kernel (espresso/modules/speech_convolutions.py:78-102).
// Activations are channels-last [B, T, F, C]; a tile of im2col rows is a (bt x bf) box of output positions, and the A
// tile of one filter tap is that box of the INPUT shifted by the tap (and strided by the convolution stride through the
// tensor map's element strides) -- one TMA box load, zero padding = TMA out-of-bounds fill.  Nothing is materialised.
//   mode 1 (fprop / dgrad): A = implicit im2col (K-major), K block kb = (tap, 64-channel block); output rows are
//           scattered back to positions (strided for the parity classes of a stride-2 dgrad);
//   mode 2 (wgrad): the reduction runs over POSITIONS (K block = one 64-position box); A = dY^T, B = the shifted input
//           box of the tap that owns the N chunk, both MN-major; the epilogue is the ordinary fp32 accumulate.
struct ConvGeom {
  int mode;
  int bf_log2, bt;        // box of positions: bf = 1 << bf_log2 along F, bt along T  (bt * bf = 128 in mode 1, 64 in mode 2)
  int tiles_t, tiles_f;   // boxes per utterance
  int st, sf;             // box start in the tensor behind the shifted loads = box index * (bt * st, bf * sf) + offset[tap]
  int cblocks;            // mode 1: 64-channel K blocks per tap
  int cin;                // mode 2: channels per tap along N
  int OT, OF, ost, osf, opt, opf;  // mode 1 output mapping: (t, f) -> (t * ost + opt, f * osf + opf), valid below (OT, OF)
  signed char offt[9], offf[9], btap[9];  // per tap: offsets of the shifted box; tap coordinate in the (ci, tap, co) weight view
};

struct KParams {
  int M, N, K, nb1, nb2, ksplit;
  ConvGeom cv;
  float* rowsum_a;     // optional: rowsum_a[m] += rowsum_scale * sum_k A[m, k] (bias gradient of a weight-gradient GEMM)
  float rowsum_scale;
  int tma_c;  // the bf16 output goes through per-warp shared-memory slabs and TMA stores (tmC is valid)
  int debug;  // ESP_GEMM_DEBUG bit mask for bottleneck experiments (0 in production): 1 no epilogue stores, 2 no MMA,
              // 4 no TMA loads, 8 no TMEM reads either
  int a_b1, a_b2, b_b1, b_b2;  // 0 => operand is broadcast along that batch dim
  EpiParams ep;
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  uint32_t spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    // A protocol bug must surface as a launch failure, never as a hung GPU.
    if (++spins > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// multicast variants for a 2-CTA cluster: data and the mbarrier signal land at the same CTA-relative offsets in every
// CTA of `mask`
__device__ __forceinline__ void tma_load_4d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1,
                                               int c2, int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      :
      : "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
// ---- TMA store (shared -> global, bulk async-group completion) -------------------------------------------------
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tm), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- cta_group::2 (CTA pair, one 256-row UMMA across two SMs) ------------------------------------------------
// Shared::cluster addresses carry the CTA rank in bit 24; clearing it makes an mbarrier address name the LEADER
// (even) CTA's barrier at the same offset (cute/arch/copy_sm100_tma.hpp Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(tm), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_cg2_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster.  RELAXED: the only thing handed over is "this
// warp's tcgen05.ld of the accumulator has completed", which tcgen05.fence::before_thread_sync orders; a release at cluster
// scope would additionally drain the warp's outstanding global stores (stall_membar was the hottest epilogue stall).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32 consecutive bf16 as four 16-byte vectors (caller guarantees alignment and range)
struct Packed32 { uint4 q[4]; };
__device__ __forceinline__ void ldg32(const bf16* p, Packed32& o) {
#pragma unroll
  for (int j = 0; j < 4; ++j) o.q[j] = __ldg(reinterpret_cast<const uint4*>(p) + j);
}
__device__ __forceinline__ void unpack32(const Packed32& o, float* v) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    unpack_bf16x2(o.q[j].x, v[8 * j + 0], v[8 * j + 1]);
    unpack_bf16x2(o.q[j].y, v[8 * j + 2], v[8 * j + 3]);
    unpack_bf16x2(o.q[j].z, v[8 * j + 4], v[8 * j + 5]);
    unpack_bf16x2(o.q[j].w, v[8 * j + 6], v[8 * j + 7]);
  }
}
__device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Shared-memory matrix descriptor (SWIZZLE_128B).  Field layout: cute/arch/mma_sm100_desc.hpp
// (SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) |
// layout_type=2 (SWIZZLE_128B) [61,64).
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes,
                                               uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// sigmoid through the single-MUFU tanh approximation (|err| ~ 5e-4: below bf16 output resolution)
__device__ __forceinline__ float fast_sigmoid(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ void load8f(const bf16* p, float* v) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  unpack_bf16x2(q.x, v[0], v[1]);
  unpack_bf16x2(q.y, v[2], v[3]);
  unpack_bf16x2(q.z, v[4], v[5]);
  unpack_bf16x2(q.w, v[6], v[7]);
}
// 32 consecutive bf16 (vectorised when 16-byte aligned and fully in range)
__device__ __forceinline__ void load32(const bf16* p, int valid, float* v) {
  if (valid >= 32 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) load8f(p + j, v + j);
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = j < valid ? bf2f(p[j]) : 0.f;
  }
}
// 32-byte (one full L2 sector) global store: sm_100 has 256-bit st.global.  A row-per-thread epilogue that writes 16 bytes
// per instruction leaves every sector half-written per request; the sub-sector writes were the dominant cost of the whole
// GEMM (profiles/r02_gemm_bottleneck.txt: 7.3 of 22.5 us).
__device__ __forceinline__ void stg256(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                       uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
               "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}
__device__ __forceinline__ void store32_bf16(bf16* p, int valid, const float* v) {
  if (valid >= 32 && ((reinterpret_cast<uintptr_t>(p) & 31) == 0)) {
#pragma unroll
    for (int j = 0; j < 32; j += 16)
      stg256(p + j, pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]), pack_bf16x2(v[j + 4], v[j + 5]),
             pack_bf16x2(v[j + 6], v[j + 7]), pack_bf16x2(v[j + 8], v[j + 9]), pack_bf16x2(v[j + 10], v[j + 11]),
             pack_bf16x2(v[j + 12], v[j + 13]), pack_bf16x2(v[j + 14], v[j + 15]));
  } else if (valid >= 32 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      uint4 o;
      o.x = pack_bf16x2(v[j], v[j + 1]);
      o.y = pack_bf16x2(v[j + 2], v[j + 3]);
      o.z = pack_bf16x2(v[j + 4], v[j + 5]);
      o.w = pack_bf16x2(v[j + 6], v[j + 7]);
      *reinterpret_cast<uint4*>(p + j) = o;
    }
  } else {
    for (int j = 0; j < 32; ++j)
      if (j < valid) p[j] = f2bf(v[j]);
  }
}
// dropout over 32 consecutive logical indices starting at idx0: one hash per 4 elements when aligned
__device__ __forceinline__ void dropout32(float* v, unsigned long long seed, unsigned long long idx0, uint32_t thresh,
                                          float scale) {
  if ((idx0 & 3ull) == 0) {
#pragma unroll
    for (int g4 = 0; g4 < 8; ++g4) {
      const unsigned long long h = esp_hash_u64(seed, (idx0 >> 2) + g4);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t lane16 = (uint32_t)(h >> (16 * t)) & 0xFFFFu;
        v[g4 * 4 + t] = lane16 >= thresh ? v[g4 * 4 + t] * scale : 0.f;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = esp_dropout_keep(seed, idx0 + j, thresh) ? v[j] * scale : 0.f;
  }
}

template <int BN, bool A_K, bool B_K, bool CG2 = false, bool TS = false>
struct SmemLayout {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (CG2 ? BN / 2 : BN) * BK * 2;  // cta_group::2: each CTA holds half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  // TS (TMA-store epilogue): one ring stage less pays for the output slabs
  static constexpr int kStages = TS ? (CG2 ? 5 : (BN == 256 ? 3 : 5)) : ((BN == 256 && !CG2) ? 4 : 6);
  // output staging for the TMA-store epilogue: TWO 32-row x 64-column bf16 slabs (SWIZZLE_128B, 4 KB each) per epilogue
  // warp, used alternately so that a warp only waits for the bulk store issued two chunk pairs ago
  static constexpr int kSlabBytes = 32 * 128;
  static constexpr int kSlabsPerWarp = TS ? 2 : 0;
  static constexpr int kSlabOffset = kStages * kStageBytes;
  static constexpr int kBarOffset = kSlabOffset + kEpiWarps * kSlabsPerWarp * kSlabBytes;
  static constexpr int kTotal = kBarOffset + 256 + 1024;  // + alignment slack
};

// MC = 2: the kernel runs as 2-CTA clusters along M.  The two CTAs of a cluster work on M-adjacent tiles of the same
// (N tile, batch, K split), so they need the SAME B tile: each loads half of it and multicasts it to both, which cuts
// the L2->SM operand traffic per CTA from (128 + BN) to (128 + BN/2) rows per k-block -- the bound of these tiles.
//
// MC = 3: cta_group::2.  The pair computes ONE 256 x BN tile with tcgen05.mma.cta_group::2 (UMMA_M = 256): each CTA
// stages its own 128 rows of A and HALF of the B tile (BN/2 rows), the leader CTA's MMA thread issues for both SMs, each
// CTA's TMEM receives the accumulator of its own 128 rows and its own epilogue warps drain it.  Per k-block and SM the
// operand ingress drops from (128 + BN) to (128 + BN/2) rows for the same 128 x BN outputs per SM -- the bound of the
// cta_group::1 kernel.  Barrier protocol (cutlass sm100 2-SM pipelines restated): both
// producers' TMA transactions signal the LEADER's full barrier (peer bit cleared), which expects the bytes of both
// CTAs; the MMA thread's commits are multicast to both CTAs' empty / accumulator-full barriers; the epilogue warps of
// both CTAs arrive on the leader's accumulator-empty barrier.
// EW: epilogue warps, 8 (two per TMEM lane quarter, every fused epilogue) or 16 (four per quarter, 640 threads, <= 102
// registers): the LIGHT epilogues only -- optional bias, ReLU / SiLU, alpha, bf16 / fp32 store or fp32 accumulate.  With
// eight warps per SM the per-chunk chain (TMEM load -> convert -> row stores) is not hidden by other warps; sixteen halve
// each warp's share of the tile.
template <int BN, bool A_K, bool B_K, int MC, bool TS, int EW>
__global__ void __launch_bounds__(128 + 32 * EW, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const KParams p) {
  constexpr bool CG2 = (MC == 3);
  constexpr int CL = MC > 1 ? 2 : 1;  // cluster size
  using L = SmemLayout<BN, A_K, B_K, CG2, TS>;
  constexpr int S = L::kStages;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  uint64_t* bars = (uint64_t*)(smem + L::kBarOffset);
  const uint32_t bar_base = smem_base + L::kBarOffset;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * S + kAccStages + s); };
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * S + 2 * kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // With MC = 2 a work item is a PAIR of M tiles: "tiles_m" counts pairs and this CTA owns tile 2*pair + rank
  // (a missing odd tile is all out-of-bounds rows: TMA zero-fills, the epilogue's row guard drops it).
  const int crank = MC > 1 ? (int)cluster_ctarank() : 0;
  const int tiles_m = ((p.M + BM - 1) / BM + CL - 1) / CL;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int nbatch = p.nb1 * p.nb2;
  const int num_kb_all = (p.K + BK - 1) / BK;
  const int kb_per = (num_kb_all + p.ksplit - 1) / p.ksplit;
  const int total_tiles = tiles_m * tiles_n * nbatch * p.ksplit;  // work items = output tiles (pairs) x K splits
  const int work0 = blockIdx.x / CL, work_stride = gridDim.x / CL;

  esp_pdl_trigger();  // the next kernel may be scheduled behind this grid's CTAs (it waits for our completion itself)
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      // MC = 2: the peer's producer also writes this slot, so both consumers free it; MC = 3: one (multicast) commit
      // (+2: the two row-sum warps also read the A tile of every stage)
      mbar_init(empty_bar(s), (MC == 2 ? 2 : 1) + ((!A_K && p.rowsum_a) ? 2 : 0));
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), CG2 ? 2 * EW : EW);  // one arrive per epilogue warp (of both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    constexpr int kCols = kAccStages * BN;  // 128 / 256 / 512: all powers of two >= 32
    if (CG2) {  // the same warp of BOTH CTAs allocates (cute::TMEM::Allocator2Sm contract)
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(kCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(kCols));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (MC > 1) cluster_sync_all();  // peer barriers are initialised before any multicast / remote arrive
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barrier init, TMEM allocation and descriptor prefetch above overlap the previous kernel's tail; nothing before
  // this point touches global memory written by earlier kernels
  esp_pdl_wait();

  if (warp == 0) {
    // ================================ TMA producer =====================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int work = work0; work < total_tiles; work += work_stride) {
        const int ks = work % p.ksplit;
        const int tile = work / p.ksplit;
        const int kb0 = ks * kb_per;
        const int kb1 = min(num_kb_all, kb0 + kb_per);
        if (kb0 >= kb1) continue;  // empty K split (all three roles skip it identically)
        const int mt = (tile % tiles_m) * CL + crank;
        const int rest = tile / tiles_m;
        const int nt = rest % tiles_n;
        const int bt = rest / tiles_n;
        const int b1 = bt % p.nb1, b2 = bt / p.nb1;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          const uint32_t sa = smem_base + s * L::kStageBytes;
          const uint32_t sb = sa + L::kABytes;
          if (p.debug & 4) {  // experiment: no operand traffic at all
            if (!CG2 || crank == 0) mbar_arrive(full_bar(s));
            continue;
          }
          if (CG2) {
            // both CTAs' transactions complete on the leader's barrier, which expects the bytes of the pair
            if (crank == 0) mbar_expect_tx(full_bar(s), 2 * L::kStageBytes);
            if (A_K) {
              tma_load_4d_cg2(sa, &tmA, full_bar(s), kb * BK, mt * BM, b1 * p.a_b1, b2 * p.a_b2);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_4d_cg2(sa + j * (BK * 128), &tmA, full_bar(s), mt * BM + 64 * j, kb * BK, b1 * p.a_b1,
                                b2 * p.a_b2);
            }
            const int n0 = nt * BN + crank * (BN / 2);  // this CTA's half of the B tile
            if (B_K) {
              tma_load_4d_cg2(sb, &tmB, full_bar(s), kb * BK, n0, b1 * p.b_b1, b2 * p.b_b2);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j)
                tma_load_4d_cg2(sb + j * (BK * 128), &tmB, full_bar(s), n0 + 64 * j, kb * BK, b1 * p.b_b1,
                                b2 * p.b_b2);
            }
            continue;
          }
          mbar_expect_tx(full_bar(s), L::kStageBytes);
          if (MC == 1 && p.cv.mode == 1) {
            // implicit im2col: the tile's box of positions, shifted by this K block's tap
            const ConvGeom& g = p.cv;
            const int tf = mt % g.tiles_f, tt = (mt / g.tiles_f) % g.tiles_t, cb = mt / (g.tiles_f * g.tiles_t);
            const int tap = kb / g.cblocks, c0 = (kb - tap * g.cblocks) * 64;
            tma_load_4d(sa, &tmA, full_bar(s), c0, (tf << g.bf_log2) * g.sf + g.offf[tap], tt * g.bt * g.st + g.offt[tap], cb);
            if (B_K) {
              tma_load_4d(sb, &tmB, full_bar(s), kb * BK, nt * BN, 0, 0);  // weights [Cout, (tap, ci)]
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)  // weights viewed (ci, tap, co): rows = 64 output channels of this tap
                tma_load_4d(sb + j * (BK * 128), &tmB, full_bar(s), nt * BN + 64 * j, g.btap[tap], c0, 0);
            }
            continue;
          }
          if (MC == 1 && p.cv.mode == 2) {
            // weight gradient: K block = one box of 64 positions; A = dY^T, B = the input box shifted by the chunk's tap
            const ConvGeom& g = p.cv;
            const int tf = kb % g.tiles_f, tt = (kb / g.tiles_f) % g.tiles_t, cb = kb / (g.tiles_f * g.tiles_t);
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_4d(sa + j * (BK * 128), &tmA, full_bar(s), mt * BM + 64 * j, tf << g.bf_log2, tt * g.bt, cb);
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) {
              const int n = nt * BN + 64 * j;
              if (n < p.N) {
                const int tap = n / g.cin, c0 = n - tap * g.cin;
                tma_load_4d(sb + j * (BK * 128), &tmB, full_bar(s), c0, (tf << g.bf_log2) * g.sf + g.offf[tap],
                            tt * g.bt * g.st + g.offt[tap], cb);
              } else {  // chunk beyond the last tap: an all-out-of-bounds box (zero fill; the bytes still arrive)
                tma_load_4d(sb + j * (BK * 128), &tmB, full_bar(s), g.cin, 0, 0, cb);
              }
            }
            continue;
          }
          if (A_K) {
            tma_load_4d(sa, &tmA, full_bar(s), kb * BK, mt * BM, b1 * p.a_b1, b2 * p.a_b2);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j)
              tma_load_4d(sa + j * (BK * 128), &tmA, full_bar(s), mt * BM + 64 * j, kb * BK,
                          b1 * p.a_b1, b2 * p.a_b2);
          }
          if (MC != 2) {
            if (B_K) {
              tma_load_4d(sb, &tmB, full_bar(s), kb * BK, nt * BN, b1 * p.b_b1, b2 * p.b_b2);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_4d(sb + j * (BK * 128), &tmB, full_bar(s), nt * BN + 64 * j, kb * BK,
                            b1 * p.b_b1, b2 * p.b_b2);
            }
          } else {
            // this CTA fetches half `crank` of the B tile for both CTAs (the tensor map's box is BN/2 rows when
            // B is K-major; MN-major tiles are made of 64-wide chunks, half of them each)
            if (B_K) {
              tma_load_4d_mc(sb + crank * (BN / 2) * 128, &tmB, full_bar(s), kb * BK, nt * BN + crank * (BN / 2),
                             b1 * p.b_b1, b2 * p.b_b2, (uint16_t)0x3);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 128; ++j) {
                const int jj = crank * (BN / 128) + j;
                tma_load_4d_mc(sb + jj * (BK * 128), &tmB, full_bar(s), nt * BN + 64 * jj, kb * BK, b1 * p.b_b1,
                               b2 * p.b_b2, (uint16_t)0x3);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer =======================================
    if (lane == 0 && (!CG2 || crank == 0)) {  // cta_group::2: the leader CTA issues for the pair
      // Instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D=f32, A=B=bf16.
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_K ? 0u : 1u) << 15) |
                             ((B_K ? 0u : 1u) << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)((CG2 ? 2 * BM : BM) >> 4) << 24);
      uint32_t it = 0, ti = 0;
      for (int work = work0; work < total_tiles; work += work_stride) {
        const int ks = work % p.ksplit;
        const int kb0 = ks * kb_per;
        const int kb1 = min(num_kb_all, kb0 + kb_per);
        if (kb0 >= kb1) continue;
        const int as = ti % kAccStages;
        const uint32_t aph = (ti / kAccStages) & 1;
        ++ti;
        mbar_wait(tempty_bar(as), aph ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          mbar_wait(full_bar(s), ph);
          tcgen05_fence_after();
          const uint32_t sa = smem_base + s * L::kStageBytes;
          const uint32_t sb = sa + L::kABytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 8-row groups are 1024 B apart (SBO); a 16-element K step is +32 B.
            // MN-major: 64-element MN chunks are BK*128 B apart (LBO); 8-k groups 1024 B apart
            //           (SBO); a 16-element K step is 16 rows * 128 B.
            const uint64_t da = A_K ? make_sdesc(sa + k * 32, 16, 1024)
                                    : make_sdesc(sa + k * 2048, BK * 128, 1024);
            const uint64_t db = B_K ? make_sdesc(sb + k * 32, 16, 1024)
                                    : make_sdesc(sb + k * 2048, BK * 128, 1024);
            if (p.debug & 2) continue;  // experiment: barriers only
            if (CG2) umma_bf16_cg2(tmem_d, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            else umma_bf16(tmem_d, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          // frees the smem slot when these MMAs retire (in both CTAs of a pair)
          if (CG2) tcgen05_commit_cg2_mc(empty_bar(s), (uint16_t)0x3);
          else if (MC > 1) tcgen05_commit_mc(empty_bar(s), (uint16_t)0x3);
          else tcgen05_commit(empty_bar(s));
        }
        // accumulator complete -> epilogue (of both CTAs in cta_group::2 mode)
        if (CG2) tcgen05_commit_cg2_mc(tfull_bar(as), (uint16_t)0x3);
        else tcgen05_commit(tfull_bar(as));
      }
    }
  } else if (!A_K && p.rowsum_a != nullptr && (warp == 2 || warp == 3)) {
    // ================================ bias gradient ======================================
    // Weight-gradient GEMMs (A = dy^T, MN-major): the row sums of A over the whole reduction are the bias gradient of the
    // same layer.  The two warps that are idle during the mainloop read every A tile from shared memory (the data the
    // tensor core consumes anyway) and accumulate it -- the separate column-sum pass over dy disappears.
    // A tile layout (MN-major, SWIZZLE_128B): BM/64 boxes of [BK k-rows][64 m]; row k of a box is 128 bytes, its 16-byte
    // unit u (8 consecutive m) lives at unit u ^ (k & 7).  Lane l owns unit (l & 15) of the 16 units of a k-row pair.
    const int u = lane & 15;
    const uint32_t unit_base = (uint32_t)(u >> 3) * (BK * 128);
    const int u7 = u & 7;
    const int parts = tiles_n >= 4 ? 4 : (tiles_n >= 2 ? 2 : 1);  // CTAs per M tile that share the summing
    const int rows_per = 32 / parts;                              // k rows per warp, CTA and k-block
    uint32_t it = 0;
    for (int work = work0; work < total_tiles; work += work_stride) {
      const int ks = work % p.ksplit;
      const int tile = work / p.ksplit;
      const int kb0 = ks * kb_per;
      const int kb1 = min(num_kb_all, kb0 + kb_per);
      if (kb0 >= kb1) continue;
      const int mt = (tile % tiles_m) * CL + crank;
      const int nt = (tile / tiles_m) % tiles_n;
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const int s = it % S;
        mbar_wait(full_bar(s), (it / S) & 1);
        if (nt < parts) {
          // The CTAs that share this M tile (one per N tile) split the k rows of every tile between them, so each of the
          // two warps reads 16 / parts row pairs per k-block (every CTA still takes part in the barrier protocol).
          const uint8_t* ta_ = smem + s * L::kStageBytes + unit_base;
          const int kbase = (warp - 2) * 32 + nt * rows_per + (lane >> 4);
          uint4 q4[8];
#pragma unroll 1
          for (int i0 = 0; i0 < rows_per / 2; i0 += 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {  // 8 independent 16-byte loads in flight
              const int k = kbase + 2 * (i0 + i);
              q4[i] = (i0 + i < rows_per / 2) ? *reinterpret_cast<const uint4*>(ta_ + k * 128 + ((u7 ^ (k & 7)) << 4))
                                              : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              // bf16 -> fp32 is a shift / mask of the packed word
              acc[0] += __uint_as_float(q4[i].x << 16); acc[1] += __uint_as_float(q4[i].x & 0xFFFF0000u);
              acc[2] += __uint_as_float(q4[i].y << 16); acc[3] += __uint_as_float(q4[i].y & 0xFFFF0000u);
              acc[4] += __uint_as_float(q4[i].z << 16); acc[5] += __uint_as_float(q4[i].z & 0xFFFF0000u);
              acc[6] += __uint_as_float(q4[i].w << 16); acc[7] += __uint_as_float(q4[i].w & 0xFFFF0000u);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
      }
      if (nt < parts) {
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
        if (lane < 16) {
          const int m0 = mt * BM + u * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (m0 + e < p.M) atomicAdd(p.rowsum_a + m0 + e, acc[e] * p.rowsum_scale);
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue =========================================
    // 8 warps: warp w reads TMEM lane quarter (w & 3); the two warps of a quarter split the tile's columns.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;  // which share of the tile's columns (2 shares with EW = 8, 4 with EW = 16)
    constexpr bool LIGHT = (EW == 16);
    constexpr int kChunks = BN / 32;          // 32-column chunks per tile
    constexpr int kChunksPerWarp = kChunks / (EW / 4);
    static_assert(kChunksPerWarp >= 1, "too many epilogue warps for this tile width");
    const EpiParams& e = p.ep;
    const unsigned long long seed = e.seed + (e.seed_ptr ? *e.seed_ptr : 0ull);
    constexpr bool use_slab = TS;
    uint8_t* const slab0 = smem + L::kSlabOffset + (warp - 4) * 2 * L::kSlabBytes;
    uint8_t* slab = slab0;
    uint32_t pair = 0;  // chunk pairs sent so far: slab (pair & 1) is the one being filled
    uint32_t ti = 0;
    for (int work = work0; work < total_tiles; work += work_stride) {
      const int ks = work % p.ksplit;
      const int tile = work / p.ksplit;
      const int kb0 = ks * kb_per;
      const int kb1 = min(num_kb_all, kb0 + kb_per);
      if (kb0 >= kb1) continue;
      const int mt = (tile % tiles_m) * CL + crank;
      const int rest = tile / tiles_m;
      const int nt = rest % tiles_n;
      const int bt = rest / tiles_n;
      const int b1 = bt % p.nb1, b2 = bt / p.nb1;
      const int as = ti % kAccStages;
      const uint32_t aph = (ti / kAccStages) & 1;
      ++ti;
      mbar_wait(tfull_bar(as), aph);
      tcgen05_fence_after();

      const int m = mt * BM + q * 32 + lane;
      bool row_ok = m < p.M;
      long c_off = (long)b1 * e.sC1 + (long)b2 * e.sC2 + (long)m * e.ldc;
      if (MC == 1 && p.cv.mode == 1) {
        // row of an im2col tile -> its output position (channels-last activation; strided for dgrad parity classes)
        const ConvGeom& g = p.cv;
        const int tf = mt % g.tiles_f, tt = (mt / g.tiles_f) % g.tiles_t, cb = mt / (g.tiles_f * g.tiles_t);
        const int ml = q * 32 + lane;
        const int ot = (tt * g.bt + (ml >> g.bf_log2)) * g.ost + g.opt;
        const int of = ((tf << g.bf_log2) + (ml & ((1 << g.bf_log2) - 1))) * g.osf + g.opf;
        row_ok = ot < g.OT && of < g.OF;
        c_off = (((long)cb * g.OT + ot) * g.OF + of) * e.ldc;
      }
      const long aux_off = (long)b1 * e.sAux1 + (long)b2 * e.sAux2 + (long)m * e.ld_aux;
      const long r_off = (long)b1 * e.sR1 + (long)b2 * e.sR2 + (long)m * e.ldr;
      const unsigned long long rng_row = ((unsigned long long)bt * p.M + m) * (unsigned long long)p.N;

#pragma unroll 1
      for (int cc = 0; cc < kChunksPerWarp; ++cc) {
        const int c = half * kChunksPerWarp + cc;
        const int n0 = nt * BN + c * 32;
        if (n0 >= p.N) break;  // warp-uniform
        if (p.debug & 8) break;  // experiment: the epilogue only hands the accumulator back
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN + c * 32), r);  // asynchronous
        const int valid = p.N - n0;  // >= 1; >= 32 for a full chunk
        if constexpr (LIGHT) {
          // light epilogue (host guarantees: no second output, aux, residual, skew or dropout)
          Packed32 pkb;
          const bf16* bp = e.bias ? e.bias + n0 : nullptr;
          const bool fb = row_ok && valid >= 32 && !e.atomic && bp && al16(bp);
          if (fb) ldg32(bp, pkb);
          tmem_ld_wait();
          if (row_ok && !(p.debug & 1)) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (e.atomic) {
              float* cp = (float*)e.C + c_off + n0;
              if (valid >= 32 && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  red_add_v4(cp + j, v[j] * e.alpha, v[j + 1] * e.alpha, v[j + 2] * e.alpha, v[j + 3] * e.alpha);
              } else {
                for (int j = 0; j < 32; ++j)
                  if (j < valid) atomicAdd(cp + j, v[j] * e.alpha);
              }
            } else {
              if (bp) {
                float bv[32];
                if (fb) unpack32(pkb, bv);
                else load32(bp, valid, bv);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += bv[j];
              }
              if (e.act == ESP_ACT_SILU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= fast_sigmoid(v[j]);
              } else if (e.act == ESP_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
              }
              if (e.alpha != 1.f) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
              }
              if (e.c_f32) {
                float* cp = (float*)e.C + c_off + n0;
                if (valid >= 32 && ((reinterpret_cast<uintptr_t>(cp) & 31) == 0)) {
#pragma unroll
                  for (int j = 0; j < 32; j += 8)
                    stg256(cp + j, __float_as_uint(v[j]), __float_as_uint(v[j + 1]), __float_as_uint(v[j + 2]),
                           __float_as_uint(v[j + 3]), __float_as_uint(v[j + 4]), __float_as_uint(v[j + 5]),
                           __float_as_uint(v[j + 6]), __float_as_uint(v[j + 7]));
                } else {
                  for (int j = 0; j < 32; ++j)
                    if (j < valid) cp[j] = v[j];
                }
              } else {
                store32_bf16((bf16*)e.C + c_off + n0, valid, v);
              }
            }
          }
          continue;
        }
        // Put every global operand of this chunk in flight BEFORE waiting for the TMEM load, so the latencies
        // (TMEM, bias from L2, aux / residual from HBM) overlap instead of adding up.
        Packed32 pk_bias, pk_aux, pk_res;
        const bf16* bias_p = e.bias ? e.bias + n0 : nullptr;
        const bf16* aux_p = (e.act >= ESP_ACT_RELU_BWD) ? e.aux + aux_off + n0 : nullptr;
        const bf16* res_p = (e.R && !e.skew_r && !e.r_f32) ? (const bf16*)e.R + r_off + n0 : nullptr;
        const bool full = row_ok && valid >= 32 && !e.atomic;
        const bool f_bias = full && bias_p && al16(bias_p);
        const bool f_aux = full && aux_p && al16(aux_p);
        const bool f_res = full && res_p && al16(res_p);
        if (f_bias) ldg32(bias_p, pk_bias);
        if (f_aux) ldg32(aux_p, pk_aux);
        if (f_res) ldg32(res_p, pk_res);
        // skewed residual (relative-position scores): 32 consecutive but arbitrarily aligned bf16 -> 17 aligned words,
        // re-aligned with a funnel shift when the first element sits in the upper half of its word
        uint32_t sk[17];
        const bool f_skew = full && e.R && e.skew_r;
        bool sk_odd = false;
        if (f_skew) {
          const bf16* rr = (const bf16*)e.R + r_off + (e.skew_r - 1 - m) + n0;
          sk_odd = (reinterpret_cast<uintptr_t>(rr) & 2) != 0;
          const uint32_t* wp = reinterpret_cast<const uint32_t*>(rr - (sk_odd ? 1 : 0));
#pragma unroll
          for (int j = 0; j < 17; ++j) sk[j] = __ldg(wp + j);
        }
        tmem_ld_wait();
        if (row_ok && !(p.debug & 1)) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (e.atomic) {
          // split-K / gradient accumulation: C (fp32) += alpha * acc, vector reductions at L2
          float* cp = (float*)e.C + c_off + n0;
          if (valid >= 32 && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(cp + j, v[j] * e.alpha, v[j + 1] * e.alpha, v[j + 2] * e.alpha, v[j + 3] * e.alpha);
          } else {
            for (int j = 0; j < 32; ++j)
              if (j < valid) atomicAdd(cp + j, v[j] * e.alpha);
          }
        } else {
        if (bias_p) {
          float bv[32];
          if (f_bias) unpack32(pk_bias, bv);
          else load32(bias_p, valid, bv);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += bv[j];
        }
        if (e.C2) store32_bf16(e.C2 + c_off + n0, valid, v);
        if (e.drop_mode == 2) dropout32(v, seed, rng_row + n0, e.drop_thresh, e.drop_scale);
        if (e.act == ESP_ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= fast_sigmoid(v[j]);
        } else if (e.act == ESP_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (e.act == ESP_ACT_SILU_BWD) {
          float u[32];
          if (f_aux) unpack32(pk_aux, u);
          else load32(aux_p, valid, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sg = fast_sigmoid(u[j]);
            v[j] *= sg * fmaf(u[j], 1.f - sg, 1.f);
          }
        } else if (e.act == ESP_ACT_RELU_BWD) {
          float u[32];
          if (f_aux) unpack32(pk_aux, u);
          else load32(aux_p, valid, u);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = u[j] > 0.f ? v[j] : 0.f;
        }
        if (e.drop_mode == 1) dropout32(v, seed, rng_row + n0, e.drop_thresh, e.drop_scale);
        if (e.alpha != 1.f) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
        }
        if (e.R) {
          if (e.skew_r) {
            // Transformer-XL skew fused into the score GEMM: R is BD_full[row m, (skew_r-1)-m+n]
            // (fairseq/modules/multihead_attention.py:824-830 as_strided trick).
            if (f_skew) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const uint32_t w = sk_odd ? __funnelshift_r(sk[j], sk[j + 1], 16) : sk[j];
                float lo, hi;
                unpack_bf16x2(w, lo, hi);
                v[2 * j] += e.beta * lo;
                v[2 * j + 1] += e.beta * hi;
              }
            } else {
              const bf16* rr = (const bf16*)e.R + r_off + (e.skew_r - 1 - m) + n0;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) v[j] += e.beta * bf2f(rr[j]);
            }
          } else if (e.r_f32) {
            const float* rr = (const float*)e.R + r_off + n0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < valid) v[j] += e.beta * rr[j];
          } else {
            float rv[32];
            if (f_res) unpack32(pk_res, rv);
            else load32(res_p, valid, rv);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(e.beta, rv[j], v[j]);
          }
        }
        if (e.c_f32) {
          float* cp = (float*)e.C + c_off + n0;
          if (valid >= 32 && ((reinterpret_cast<uintptr_t>(cp) & 31) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              stg256(cp + j, __float_as_uint(v[j]), __float_as_uint(v[j + 1]), __float_as_uint(v[j + 2]),
                     __float_as_uint(v[j + 3]), __float_as_uint(v[j + 4]), __float_as_uint(v[j + 5]),
                     __float_as_uint(v[j + 6]), __float_as_uint(v[j + 7]));
          } else if (valid >= 32 && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (j < valid) cp[j] = v[j];
          }
        } else if (use_slab) {
          // this row's 64 bytes of the chunk -> the warp's SWIZZLE_128B slab (16-byte unit u of row `lane` lives at
          // unit u ^ (lane & 7)); the pair of chunks leaves as ONE bulk tensor store of full 128-byte lines
          uint8_t* srow = slab + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_bf16x2(v[8 * j], v[8 * j + 1]);
            o.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
            o.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
            o.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
            *reinterpret_cast<uint4*>(srow + ((((cc & 1) * 4 + j) ^ (lane & 7)) << 4)) = o;
          }
        } else {
          store32_bf16((bf16*)e.C + c_off + n0, valid, v);
        }
        }  // !atomic
        }  // row_ok
        if (use_slab && ((cc & 1) || cc == kChunksPerWarp - 1 || n0 + 32 >= p.N)) {
          // hand the slab to the TMA engine (rows >= M and columns >= N are clipped by the tensor map) and move on to the
          // other slab: it is free once the store issued from it two pairs ago has been READ (at most one group pending)
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmC, smem_u32(slab), nt * BN + (half * kChunksPerWarp + (cc & ~1)) * 32, mt * BM + q * 32, b1, b2);
            tma_store_commit();
            tma_store_wait_read1();
          }
          ++pair;
          slab = slab0 + (pair & 1) * L::kSlabBytes;
          __syncwarp();
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG2) mbar_arrive_cluster(tempty_bar(as), 0u);  // the leader's MMA thread waits for both CTAs' epilogues
        else mbar_arrive(tempty_bar(as));
      }
    }
    if (use_slab && lane == 0) tma_store_wait_all();  // every bulk store of this warp has been written
  }

  tcgen05_fence_before();
  __syncthreads();
  if (MC > 1) cluster_sync_all();  // no CTA exits while its peer may still signal or write into its shared memory
  if (warp == 2) {
    tcgen05_fence_after();
    constexpr int kCols = kAccStages * BN;
    if (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kCols));
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
bool esp_gemm_multicast_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("ESP_GEMM_MULTICAST");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// rows x inner (inner contiguous), two batch dims.  box = {box_inner(64), box_rows}.
int make_tmap(CUtensorMap* tm, const void* base, long inner, long rows, long ld, int nb1, long s1,
              int nb2, long s2, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  ESP_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  // cuTensorMapEncodeTiled is a DRIVER call: the calling thread needs a current context.  PyTorch's
  // autograd engine runs backward on its own threads where only the runtime device is set, so bind the
  // primary context once per thread (cudaFree(0) is the canonical no-op that does it).
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    ESP_CUDA(cudaFree(0));
    ctx_bound = true;
  }
  ESP_CHECK(((uintptr_t)base & 15) == 0, "GEMM operand base must be 16-byte aligned");
  ESP_CHECK((ld % 8) == 0, "GEMM operand leading dimension (%ld) must be a multiple of 8 elements", ld);
  // Broadcast / singleton batch dims get size 1 (coordinate forced to 0 by the kernel) and a
  // dummy legal stride.
  long dummy = ((rows * ld + 7) / 8) * 8;
  if (nb1 <= 1 || s1 == 0) { nb1 = 1; s1 = dummy; }
  if (nb2 <= 1 || s2 == 0) { nb2 = 1; s2 = dummy * (nb1 > 1 ? nb1 : 1); }
  ESP_CHECK((s1 % 8) == 0 && (s2 % 8) == 0, "GEMM batch strides must be multiples of 8 elements");
  cuuint64_t dims[4] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)nb1, (cuuint64_t)nb2};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)s1 * 2, (cuuint64_t)s2 * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box,
                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ESP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d): inner=%ld rows=%ld ld=%ld", (int)r,
            inner, rows, ld);
  return 0;
}

bool esp_gemm_cg2_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("ESP_GEMM_CG2");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// co-resident 2-CTA clusters of a kernel variant (clusters sit inside one GPC: with odd SM counts per GPC fewer than
// SMs/2 pairs fit at once -- ask the runtime instead of assuming)
template <typename K>
int cluster_slots(K kfn, int smem_bytes, int threads) {
  int slots = esp_num_sms() / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(esp_num_sms() / 2 * 2);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kfn, &cfg) == cudaSuccess && n > 0 && n < slots) slots = n;
  (void)cudaGetLastError();
  return slots;
}

template <int BN, bool A_K, bool B_K, int MC, bool TS = false, int EW = 8>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const KParams& kp, cudaStream_t st) {
  constexpr int CL = MC > 1 ? 2 : 1;
  using L = SmemLayout<BN, A_K, B_K, MC == 3, TS>;
  static bool configured = false;
  auto kfn = gemm_tcgen05_kernel<BN, A_K, B_K, MC, TS, EW>;
  constexpr int kT = 128 + 32 * EW;
  if (!configured) {
    ESP_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  const int tiles_m = (((kp.M + BM - 1) / BM) + CL - 1) / CL;
  const int work = tiles_m * ((kp.N + BN - 1) / BN) * kp.nb1 * kp.nb2 * kp.ksplit;
  static int slots = 0;  // persistent grid = what is co-resident
  if (slots == 0) slots = CL > 1 ? cluster_slots(kfn, L::kTotal, kT) : esp_num_sms();
  int grid = (work < slots ? work : slots) * CL;
  if (grid < 1) return 0;
  esp_launch_cluster(kfn, grid, kT, L::kTotal, st, CL, ta, tb, tc, kp);
  ESP_LAUNCH_CHECK();
  esp_count_launch(1);
  return 0;
}

template <int BN, int MC, bool TS = false, int EW = 8>
int dispatch_major(bool ak, bool bk, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
                   const KParams& kp, cudaStream_t st) {
  if (ak && bk) return launch<BN, true, true, MC, TS, EW>(ta, tb, tc, kp, st);
  if (ak && !bk) return launch<BN, true, false, MC, TS, EW>(ta, tb, tc, kp, st);
  if (!ak && bk) return launch<BN, false, true, MC, TS, EW>(ta, tb, tc, kp, st);
  return launch<BN, false, false, MC, TS, EW>(ta, tb, tc, kp, st);
}

}  // namespace

// bf16 SWIZZLE_128B tensor map (64-element boxes) for the other tcgen05 kernels of the library (attn_fused.cu)
int esp_make_tmap_bf16(CUtensorMap* tm, const void* base, long inner, long rows, long ld, int nb1, long s1, int nb2,
                       long s2, int box_rows) {
  return make_tmap(tm, base, inner, rows, ld, nb1, s1, nb2, s2, box_rows);
}

extern "C" int esp_gemm_bf16(const EspGemm* g, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ESP_CHECK(g != nullptr, "null gemm descriptor");
  ESP_CHECK(g->M >= 0 && g->N >= 0 && g->K > 0, "bad GEMM shape %ld x %ld x %ld", (long)g->M, (long)g->N,
            (long)g->K);
  if (g->M == 0 || g->N == 0) return 0;
  const int nb1 = g->nb1 > 0 ? g->nb1 : 1, nb2 = g->nb2 > 0 ? g->nb2 : 1;
  CUtensorMap ta, tb;
  const bool ak = g->a_kmajor != 0, bk = g->b_kmajor != 0;
  // tile width: wide tiles have the best operand-bytes per flop (a 128 x BN tile reads (128+BN)*64*2 B from L2 per
  // 128*BN*64*2 flop); narrow ones only when N is small or the machine would be badly under-filled.
  const int sms = esp_num_sms();
  const long tm = (g->M + BM - 1) / BM;
  auto tiles_for = [&](int w) { return tm * ((g->N + w - 1) / w) * (long)nb1 * nb2; };
  int bn;
  int ksplit = 1;
  int mode = 1;  // 1: single CTA, 2: multicast pair, 3: cta_group::2 pair
  const int num_kb = (int)((g->K + BK - 1) / BK);
  const bool cg2_ok = esp_gemm_cg2_enabled() && tm >= 2 && g->N > 128 && g->K >= 256;  // K = 64 attention tiles are epilogue-bound: no gain
  const int pairs = sms / 2;  // upper bound of co-resident clusters (the launch clamps to the real number)
  // utilisation of `slots` persistent workers by `items` equal work items
  auto eff = [](long items, long slots) {
    const long waves = (items + slots - 1) / slots;
    return (double)items / (double)(waves * slots);
  };
  if (g->accumulate) {
    // gradient GEMMs: long reduction (K = rows of the batch), small output -> split K across CTAs and
    // accumulate with vector reductions; keep >= 4 k-blocks per split.
    bn = g->N >= 192 ? 256 : (g->N > 64 ? 128 : 64);
    // (the bias-gradient warps wait on the CTA's own full barrier: not available to the peer CTA of a cta_group::2 pair)
    if (bn == 256 && cg2_ok && g->rowsum_a == nullptr) mode = 3;
    const long t = mode == 3 ? ((tm + 1) / 2) * ((g->N + 255) / 256) * (long)nb1 * nb2 : tiles_for(bn);
    const long slots = mode == 3 ? pairs : sms;
    const int max_split = num_kb / 4 > 0 ? num_kb / 4 : 1;
    double best = -1.0;
    for (int ks = 1; ks <= max_split && (long)(ks - 1) * t < slots; ++ks) {
      const double e = eff(t * ks, slots) - 0.01 * ks;  // prefer fewer splits at equal utilisation (less red traffic)
      if (e > best) { best = e; ksplit = ks; }
    }
  } else {
    if (g->N <= 64) bn = 64;
    else if (g->N >= 192 && tiles_for(256) * 10 >= (long)sms * 8) bn = 256;
    else if (g->N > 64 && tiles_for(128) * 10 >= (long)sms * 6) bn = 128;
    else if (g->N >= 192 && tiles_for(256) * 2 >= tiles_for(128)) bn = 128;
    else bn = g->N > 128 ? 128 : (g->N > 64 && tiles_for(64) < sms ? 64 : 128);
    // cta_group::2 (256 x 256 per CTA pair): whenever the output is wide enough and the pairs fill at least ~60 % of
    // the machine -- halving the per-SM operand ingress outweighs some tile-quantisation loss
    if (cg2_ok && g->N >= 192) {
      const long t2 = ((tm + 1) / 2) * ((g->N + 255) / 256) * (long)nb1 * nb2;
      if (eff(t2, pairs) >= 0.6 || t2 >= 4 * pairs) { bn = 256; mode = 3; }
    }
  }
  if (g->tile_n == 64 || g->tile_n == 128 || g->tile_n == 256) { bn = g->tile_n; mode = 1; }
  if (g->tile_n == 512) {  // forced cta_group::2 (tests, microbenchmarks)
    ESP_CHECK(tm >= 1 && g->N >= 1, "bad shape");
    bn = 256;
    mode = g->rowsum_a ? 1 : 3;
  }
  // 2-CTA multicast pairs along M: worth it when the pairing wastes (almost) no tile
  if (mode == 1 && bn >= 128 && esp_gemm_multicast_enabled() && (tm % 2 == 0 || tm >= 16)) mode = 2;
  const int mc = mode == 1 ? 1 : 2;
  int rc;
  if (ak) rc = make_tmap(&ta, g->A, g->K, g->M, g->lda, nb1, g->sA1, nb2, g->sA2, BM);
  else    rc = make_tmap(&ta, g->A, g->M, g->K, g->lda, nb1, g->sA1, nb2, g->sA2, BK);
  if (rc) return rc;
  if (bk) rc = make_tmap(&tb, g->B, g->K, g->N, g->ldb, nb1, g->sB1, nb2, g->sB2, bn / mc);
  else    rc = make_tmap(&tb, g->B, g->N, g->K, g->ldb, nb1, g->sB1, nb2, g->sB2, BK);
  if (rc) return rc;

  KParams kp;
  kp.cv.mode = 0;
  kp.rowsum_a = g->rowsum_a;
  kp.rowsum_scale = g->rowsum_scale;
  if (g->rowsum_a) {
    ESP_CHECK(!ak && g->accumulate && nb1 == 1 && nb2 == 1, "rowsum_a needs an MN-major A operand, accumulate = 1, no batch dims");
  }
  {
    const char* dbg = getenv("ESP_GEMM_DEBUG");
    kp.debug = dbg ? atoi(dbg) : 0;
  }
  kp.M = (int)g->M; kp.N = (int)g->N; kp.K = (int)g->K; kp.nb1 = nb1; kp.nb2 = nb2; kp.ksplit = ksplit;
  kp.a_b1 = (nb1 > 1 && g->sA1 != 0) ? 1 : 0;
  kp.a_b2 = (nb2 > 1 && g->sA2 != 0) ? 1 : 0;
  kp.b_b1 = (nb1 > 1 && g->sB1 != 0) ? 1 : 0;
  kp.b_b2 = (nb2 > 1 && g->sB2 != 0) ? 1 : 0;
  EpiParams& e = kp.ep;
  e.C = g->C; e.C2 = (bf16*)g->C2; e.bias = (const bf16*)g->bias; e.aux = (const bf16*)g->aux; e.R = g->R;
  e.ldc = g->ldc; e.ld_aux = g->ld_aux; e.ldr = g->ldr;
  e.sC1 = g->sC1; e.sC2 = g->sC2; e.sAux1 = g->sAux1; e.sAux2 = g->sAux2; e.sR1 = g->sR1; e.sR2 = g->sR2;
  e.c_f32 = g->c_f32; e.r_f32 = g->r_f32; e.act = g->act; e.skew_r = g->skew_r; e.atomic = g->accumulate ? 1 : 0;
  e.drop_mode = (g->drop_p > 0.f) ? g->drop_mode : 0;
  e.alpha = g->alpha; e.beta = g->beta;
  e.drop_thresh = esp_dropout_thresh(g->drop_p);
  e.drop_scale = esp_dropout_scale(g->drop_p);
  e.seed = g->seed;
  e.seed_ptr = (const unsigned long long*)g->seed_ptr;
  ESP_CHECK(g->C != nullptr, "GEMM output pointer is null");
  ESP_CHECK(!g->accumulate || g->c_f32, "accumulate (atomic) output must be fp32");
  ESP_CHECK(!(e.act >= ESP_ACT_RELU_BWD) || e.aux != nullptr, "activation-gradient epilogue needs aux");
  // Optional (ESP_GEMM_TMA_STORE=1): bf16 outputs of the wide-tile variants leave through shared-memory slabs + TMA stores
  // (full 128-byte lines, clipped at the matrix edges by the tensor map).  Correct (the whole GPU suite passes with it) but
  // measured SLOWER than the 32-byte row stores on the bench shapes (profiles/r02_gemm_microbench_v4.txt: FFN1 20.9 vs
  // 19.6 us; step 23.1 vs 22.5 ms): the epilogue is bound by TMEM read-out and per-element math, not by the store
  // pattern, and the slab hand-off adds a serial wait per chunk pair.  Kept as an experiment switch, off by default.
  CUtensorMap tc = ta;
  kp.tma_c = 0;
  {
    static int tma_on = -1;
    if (tma_on < 0) {
      const char* ev = getenv("ESP_GEMM_TMA_STORE");
      tma_on = (ev && ev[0] == '1') ? 1 : 0;
    }
    const bool ok = tma_on && bn == 256 && mode == 3 && !g->c_f32 && !g->accumulate && ((uintptr_t)g->C & 15) == 0 && g->ldc % 8 == 0 &&
                    (nb1 <= 1 || g->sC1 % 8 == 0) && (nb2 <= 1 || g->sC2 % 8 == 0) && (nb1 <= 1 || g->sC1 != 0) &&
                    (nb2 <= 1 || g->sC2 != 0);
    if (ok) {
      rc = make_tmap(&tc, g->C, g->N, g->M, g->ldc, nb1, g->sC1, nb2, g->sC2, 32);
      if (rc) return rc;
      kp.tma_c = 1;
    }
  }
  // Experiment switch ESP_GEMM_EPI16=1: sixteen epilogue warps for the light epilogues of the wide tiles.  Measured: no
  // gain (FFN1 19.6 -> 19.3 us, out_proj 11.6 -> 14.3 us; profiles/r02_gemm_microbench_v6_epi16.txt) -- the epilogue is
  // bound by the throughput of its row-per-thread stores, not by the latency of each warp's chain.  Off by default.
  static int epi16 = -1;
  if (epi16 < 0) {
    const char* ev = getenv("ESP_GEMM_EPI16");
    epi16 = (ev && ev[0] == '1') ? 1 : 0;
  }
  const bool light = epi16 && bn >= 128 && !kp.tma_c && e.C2 == nullptr && e.R == nullptr && e.drop_mode == 0 &&
                     e.act < ESP_ACT_RELU_BWD && kp.debug == 0;
  if (light) {
    if (bn == 256) {
      if (mode == 3) return dispatch_major<256, 3, false, 16>(ak, bk, ta, tb, tc, kp, st);
      return mode == 2 ? dispatch_major<256, 2, false, 16>(ak, bk, ta, tb, tc, kp, st)
                       : dispatch_major<256, 1, false, 16>(ak, bk, ta, tb, tc, kp, st);
    }
    return mode == 2 ? dispatch_major<128, 2, false, 16>(ak, bk, ta, tb, tc, kp, st)
                     : dispatch_major<128, 1, false, 16>(ak, bk, ta, tb, tc, kp, st);
  }
  if (bn == 64) return dispatch_major<64, 1>(ak, bk, ta, tb, tc, kp, st);
  if (bn == 256) {
    if (mode == 3 && kp.tma_c) return dispatch_major<256, 3, true>(ak, bk, ta, tb, tc, kp, st);
    if (mode == 3) return dispatch_major<256, 3>(ak, bk, ta, tb, tc, kp, st);
    return mode == 2 ? dispatch_major<256, 2>(ak, bk, ta, tb, tc, kp, st) : dispatch_major<256, 1>(ak, bk, ta, tb, tc, kp, st);
  }
  return mode == 2 ? dispatch_major<128, 2>(ak, bk, ta, tb, tc, kp, st) : dispatch_major<128, 1>(ak, bk, ta, tb, tc, kp, st);
}

// ------------------------------------------------------------------------------------------
// 3x3 convolutions of the conv front end (implicit GEMMs, ConvGeom above)
// ------------------------------------------------------------------------------------------
namespace {

// 4-D bf16 tensor map over a channels-last activation [B, T, F, C] (dims innermost first: C, F, T, B) or any other 4-D
// view; box[0] = 64 channels (one SWIZZLE_128B row), element strides > 1 make the box skip positions (strided convolution:
// boxDim = elements spanned, the engine writes ceil(boxDim / stride) of them densely).
int make_tmap_box(CUtensorMap* tm, const void* base, const cuuint64_t dims[4], const cuuint64_t strides_b[3],
                  const cuuint32_t box[4], const cuuint32_t estr[4]) {
  PFN_encodeTiled enc = get_encode();
  ESP_CHECK(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    ESP_CUDA(cudaFree(0));
    ctx_bound = true;
  }
  ESP_CHECK(((uintptr_t)base & 15) == 0, "conv operand base must be 16-byte aligned");
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides_b, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ESP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (conv box) failed (%d): dims %llu %llu %llu %llu box %u %u %u %u", (int)r,
            (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2], (unsigned long long)dims[3],
            box[0], box[1], box[2], box[3]);
  return 0;
}

int act_tmap(CUtensorMap* tm, const void* base, int B, int T, int F, int C, int box_f, int box_t, int sf, int st) {
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)F * C * 2, (cuuint64_t)T * F * C * 2};
  const cuuint32_t box[4] = {64, (cuuint32_t)(box_f * sf), (cuuint32_t)(box_t * st), 1};
  const cuuint32_t estr[4] = {1, (cuuint32_t)sf, (cuuint32_t)st, 1};
  return make_tmap_box(tm, base, dims, strides, box, estr);
}

// box of `n` positions: as wide along F as divides it evenly (power of two, at most 16), the rest along T
void pick_box(int n, int Fpos, int* bf_log2, int* bt) {
  int l = 0;
  while (l < 4 && (1 << (l + 1)) <= n && Fpos % (1 << (l + 1)) == 0) ++l;
  *bf_log2 = l;
  *bt = n >> l;
}

void zero_epi(KParams& kp) {
  EpiParams& e = kp.ep;
  e.C = nullptr; e.C2 = nullptr; e.bias = nullptr; e.aux = nullptr; e.R = nullptr;
  e.ldc = e.ld_aux = e.ldr = 0;
  e.sC1 = e.sC2 = e.sAux1 = e.sAux2 = e.sR1 = e.sR2 = 0;
  e.c_f32 = e.r_f32 = e.act = e.drop_mode = e.skew_r = e.atomic = 0;
  e.alpha = 1.f; e.beta = 0.f; e.drop_scale = 1.f; e.drop_thresh = 0; e.seed = 0; e.seed_ptr = nullptr;
  kp.rowsum_a = nullptr; kp.rowsum_scale = 0.f; kp.tma_c = 0; kp.debug = 0;
  kp.nb1 = kp.nb2 = 1; kp.ksplit = 1; kp.a_b1 = kp.a_b2 = kp.b_b1 = kp.b_b2 = 0;
}

int conv_check(int B, int T, int F, int Cin, int Cout, int st, int sf) {
  ESP_CHECK(B > 0 && T > 0 && F > 0, "conv3x3: empty input");
  ESP_CHECK(Cin % 64 == 0 && Cout % 64 == 0, "conv3x3 (tensor-core path): channel counts must be multiples of 64 (got %d -> %d)", Cin, Cout);
  ESP_CHECK((st == 1 || st == 2) && (sf == 1 || sf == 2), "conv3x3: strides 1 or 2 (got %d x %d)", st, sf);
  return 0;
}

}  // namespace

// y[b, to, fo, :] = sum_{r,s,ci} x[b, to*st + r - 1, fo*sf + s - 1, ci] * w[:, r, s, ci]      (zero padding 1, no bias)
// x [B, T, F, Cin], w [Cout, 3, 3, Cin], y [B, ceil(T/st), ceil(F/sf), Cout], all bf16 channels-last.
extern "C" int esp_conv3x3_fwd(const void* x, const void* w, void* y, int32_t B, int32_t T, int32_t F, int32_t Cin,
                               int32_t Cout, int32_t st, int32_t sf, void* stream) {
  cudaStream_t s_ = (cudaStream_t)stream;
  if (int rc = conv_check(B, T, F, Cin, Cout, st, sf)) return rc;
  const int To = (T + st - 1) / st, Fo = (F + sf - 1) / sf;
  KParams kp;
  zero_epi(kp);
  ConvGeom& g = kp.cv;
  g.mode = 1;
  pick_box(128, Fo, &g.bf_log2, &g.bt);
  const int bf = 1 << g.bf_log2;
  g.tiles_t = (To + g.bt - 1) / g.bt;
  g.tiles_f = (Fo + bf - 1) / bf;
  g.st = st; g.sf = sf; g.cblocks = Cin / 64; g.cin = Cin;
  g.OT = To; g.OF = Fo; g.ost = g.osf = 1; g.opt = g.opf = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      g.offt[r * 3 + c] = (signed char)(r - 1);
      g.offf[r * 3 + c] = (signed char)(c - 1);
      g.btap[r * 3 + c] = (signed char)(r * 3 + c);
    }
  const int bn = Cout % 128 == 0 ? 128 : 64;
  CUtensorMap ta, tb;
  if (int rc = act_tmap(&ta, x, B, T, F, Cin, bf, g.bt, sf, st)) return rc;
  if (int rc = make_tmap(&tb, w, 9L * Cin, Cout, 9L * Cin, 1, 0, 1, 0, bn)) return rc;
  kp.M = B * g.tiles_t * g.tiles_f * BM; kp.N = Cout; kp.K = 9 * Cin;
  kp.ep.C = y; kp.ep.ldc = Cout;
  return bn == 128 ? launch<128, true, true, 1>(ta, tb, ta, kp, s_) : launch<64, true, true, 1>(ta, tb, ta, kp, s_);
}

// dx[b, t, f, :] = sum over the taps (r, s) and output positions with to*st + r - 1 == t, fo*sf + s - 1 == f of
//                  dy[b, to, fo, :] . w[:, r, s, :]
// One launch per parity class (t mod st, f mod sf): inside a class every position sees the same taps, so the class is a
// small un-strided convolution of dy whose outputs are written st x sf apart.
extern "C" int esp_conv3x3_dgrad(const void* dy, const void* w, void* dx, int32_t B, int32_t T, int32_t F, int32_t Cin,
                                 int32_t Cout, int32_t st, int32_t sf, void* stream) {
  cudaStream_t s_ = (cudaStream_t)stream;
  if (int rc = conv_check(B, T, F, Cin, Cout, st, sf)) return rc;
  const int To = (T + st - 1) / st, Fo = (F + sf - 1) / sf;
  const int bn = Cin % 128 == 0 ? 128 : 64;
  CUtensorMap tb;
  {
    // weights [Cout][tap][Cin] viewed (ci, tap, co): a K block = 64 output channels of one tap, MN-major rows of 64 ci
    const cuuint64_t dims[4] = {(cuuint64_t)Cin, 9, (cuuint64_t)Cout, 1};
    const cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)9 * Cin * 2, (cuuint64_t)9 * Cin * Cout * 2};
    const cuuint32_t box[4] = {64, 1, 64, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (int rc = make_tmap_box(&tb, w, dims, strides, box, estr)) return rc;
  }
  for (int pt = 0; pt < st; ++pt)
    for (int pf = 0; pf < sf; ++pf) {
      const int U = (T - pt + st - 1) / st, V = (F - pf + sf - 1) / sf;  // positions of this class
      if (U <= 0 || V <= 0) continue;
      KParams kp;
      zero_epi(kp);
      ConvGeom& g = kp.cv;
      g.mode = 1;
      pick_box(128, V, &g.bf_log2, &g.bt);
      const int bf = 1 << g.bf_log2;
      g.tiles_t = (U + g.bt - 1) / g.bt;
      g.tiles_f = (V + bf - 1) / bf;
      g.st = g.sf = 1; g.cblocks = Cout / 64; g.cin = Cin;
      g.OT = T; g.OF = F; g.ost = st; g.osf = sf; g.opt = pt; g.opf = pf;
      int nt = 0;
      for (int r = 0; r < 3; ++r) {
        if ((pt + 1 - r) % st != 0) continue;  // t = to*st + r - 1  =>  to = (t + 1 - r) / st = u + (pt + 1 - r) / st
        for (int c = 0; c < 3; ++c) {
          if ((pf + 1 - c) % sf != 0) continue;
          g.offt[nt] = (signed char)((pt + 1 - r) / st);
          g.offf[nt] = (signed char)((pf + 1 - c) / sf);
          g.btap[nt] = (signed char)(r * 3 + c);
          ++nt;
        }
      }
      ESP_CHECK(nt > 0, "conv3x3 dgrad: parity class without taps");
      CUtensorMap ta;
      if (int rc = act_tmap(&ta, dy, B, To, Fo, Cout, bf, g.bt, 1, 1)) return rc;
      kp.M = B * g.tiles_t * g.tiles_f * BM; kp.N = Cin; kp.K = nt * Cout;
      kp.ep.C = dx; kp.ep.ldc = Cin;
      const int rc = bn == 128 ? launch<128, true, false, 1>(ta, tb, ta, kp, s_) : launch<64, true, false, 1>(ta, tb, ta, kp, s_);
      if (rc) return rc;
    }
  return 0;
}

// dw[co, r, s, ci] += sum_{b, to, fo} dy[b, to, fo, co] * x[b, to*st + r - 1, fo*sf + s - 1, ci]      (fp32 accumulate)
extern "C" int esp_conv3x3_wgrad(const void* dy, const void* x, float* dw, int32_t B, int32_t T, int32_t F, int32_t Cin,
                                 int32_t Cout, int32_t st, int32_t sf, void* stream) {
  cudaStream_t s_ = (cudaStream_t)stream;
  if (int rc = conv_check(B, T, F, Cin, Cout, st, sf)) return rc;
  const int To = (T + st - 1) / st, Fo = (F + sf - 1) / sf;
  KParams kp;
  zero_epi(kp);
  ConvGeom& g = kp.cv;
  g.mode = 2;
  pick_box(64, Fo, &g.bf_log2, &g.bt);
  const int bf = 1 << g.bf_log2;
  g.tiles_t = (To + g.bt - 1) / g.bt;
  g.tiles_f = (Fo + bf - 1) / bf;
  g.st = st; g.sf = sf; g.cblocks = Cin / 64; g.cin = Cin;
  g.OT = To; g.OF = Fo; g.ost = g.osf = 1; g.opt = g.opf = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      g.offt[r * 3 + c] = (signed char)(r - 1);
      g.offf[r * 3 + c] = (signed char)(c - 1);
      g.btap[r * 3 + c] = (signed char)(r * 3 + c);
    }
  CUtensorMap ta, tb;
  if (int rc = act_tmap(&ta, dy, B, To, Fo, Cout, bf, g.bt, 1, 1)) return rc;
  if (int rc = act_tmap(&tb, x, B, T, F, Cin, bf, g.bt, sf, st)) return rc;
  const long kblocks = (long)B * g.tiles_t * g.tiles_f;
  kp.M = Cout; kp.N = 9 * Cin; kp.K = (int)(kblocks * BK);
  ESP_CHECK(kblocks * BK < (1L << 31), "conv3x3 wgrad: too many positions");
  const int tiles = ((Cout + BM - 1) / BM) * ((9 * Cin + 127) / 128);
  int ks = esp_num_sms() / tiles;  // fill the machine; every split keeps a long reduction (>= 8 position boxes)
  if (ks > kblocks / 8) ks = (int)(kblocks / 8);
  if (ks < 1) ks = 1;
  kp.ksplit = ks;
  kp.ep.C = dw; kp.ep.ldc = 9 * Cin; kp.ep.c_f32 = 1; kp.ep.atomic = 1;
  return launch<128, false, false, 1>(ta, tb, ta, kp, s_);
}
