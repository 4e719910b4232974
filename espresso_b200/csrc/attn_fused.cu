// espresso_b200 -- fused relative-position self-attention forward for sm_100a.
//
//   ctx[b, i, h] = sum_j dropout( softmax_j( qu_i . k_j + qv_i . p_{(T-1)-i+j}  (+ key-padding mask) ) )[i, j] v_j
//
// Reference (rel-pos branch of fairseq/modules/multihead_attention.py): scores AC = q_u k^T (:788), BD = q_v pos^T over
// the 2T-1 relative positions (:815-823), the Transformer-XL skew BD[i, j] <- BD_full[i, (T-1)-i+j] (:824-830), key-padding
// mask -> -inf (:841-859), fp32 softmax cast back to the model dtype (:864-867), dropout, P v (:897).  The reference (and
// the round-1 build) round-trips BD_full, the scores, the probabilities and the dropped probabilities through HBM; here
// the scores and the relative-position logits never leave the SM:
//
//   * one CTA per (128-query tile, head, utterance); key tiles of 128 are streamed with TMA (SWIZZLE_128B) through a
//     2-stage ring (K, V) and a 3-slot ring of 128-row relative-position blocks;
//   * per key tile three tcgen05.mma groups write TMEM: S = Qu K^T (128 columns) and W = Qv [Pblk_lo; Pblk_hi]^T
//     (2 x 128 columns) -- the 255 relative positions this (query tile, key tile) pair can see;
//   * eight softmax warps (two threads per query row = TMEM lane, each owning 64 of the 128 keys) read S with
//     tcgen05.ld, stage the row's slice of W in shared memory (bf16, as the reference keeps BD) and read it back shifted
//     by (127 - row): BD[r, c] = W[r, 127 - r + c] -- the skew is a column re-index through a per-row staging buffer, the
//     only synchronisation is a 64-thread named barrier between the two warps that share a lane quarter; S / W are
//     handed back to the tensor core as soon as the logits sit in registers, so the MMAs of tile i+1 overlap the
//     exponentials of tile i;
//   * two passes over the key tiles: pass 1 accumulates the row maximum and the normaliser online, pass 2 recomputes the
//     logits, forms the NORMALISED probabilities in registers, writes them (bf16) for the backward pass, applies the
//     counter-RNG dropout (same stream as esp_attn_softmax_fwd/bwd), stores the dropped tile as the K-major A operand in
//     shared memory and accumulates O += P V in TMEM.  Normalising before P V removes the online-softmax rescale of O.
//
// Outputs: ctx [B*T, d] bf16; when `p_out` is given (training) the probabilities P [H,B,T,ld] and, with dropout, the
// dropped probabilities Pd -- exactly what esp_attn_softmax_fwd produced, so the existing backward kernels are unchanged.
#include "common.cuh"
#include "espresso_b200.h"
#include <cuda.h>
#include <stdlib.h>

void esp_count_launch(int n);

namespace {

constexpr int kTile = 128;      // query rows per CTA = key columns per step
constexpr int kHd = 64;         // head dim (one SWIZZLE_128B row of bf16)
constexpr int kThreads = 288;   // warps 0-7: softmax (TMEM lane quarter x column half), warp 8: TMA + MMA issue + TMEM alloc
constexpr int kTileBytes = kTile * kHd * 2;  // 16 KB
constexpr int kWRowBytes = 336;              // 160 bf16 + 16 B pad (bank-conflict-free 16-byte row stores)
constexpr int kOffQu = 0;
constexpr int kOffQv = kOffQu + kTileBytes;
constexpr int kOffK = kOffQv + kTileBytes;         // 2 stages
constexpr int kOffV = kOffK + 2 * kTileBytes;      // 2 stages
constexpr int kOffP = kOffV + 2 * kTileBytes;      // 3 relative-position blocks
constexpr int kOffA = kOffP + 3 * kTileBytes;      // dropped probabilities, A operand of P V: 128 x 128 bf16 = 2 chunks
constexpr int kOffW = kOffA + 2 * kTileBytes;      // W staging, 128 rows x 336 B
constexpr int kOffBar = kOffW + kTile * kWRowBytes;
constexpr int kSmemBytes = kOffBar + 128 + 2 * kTile * 2 * 4 + 1024;   // barriers, (m, l) exchange, alignment slack
constexpr int kColS = 0, kColW = 128, kColO = 384, kTmemCols = 512;

// optional timeline of CTA (0, 0, 0) (ESP_ATTN_FWD_TIMELINE=1; esp_attn_fwd_timeline reads it): %globaltimer stamps of
// softmax warp 0 -- per iteration: logits ready (S / W in TMEM), logits in registers, iteration done
__device__ unsigned long long g_fwd_timeline[128];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TLF(slot)                                                                                                        \
  do {                                                                                                                   \
    if (p.timeline && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (slot) < 128)         \
      g_fwd_timeline[(slot)] = gtimer();                                                                                 \
  } while (0)

struct Params {
  int timeline;
  int B, T, H, d, ld;       // ld: row stride of the probability tensors (multiple of 8, >= T)
  int pos_hstride;          // head stride (elements) inside a projected-position row: hd, or 0 (table shared by heads)
  const int* lens;          // valid keys per utterance or nullptr
  const int* key_lo;        // optional per-query-row visible key range [key_lo[i], key_hi[i]) (streaming / context masks)
  const int* key_hi;
  bf16* ctx;                // [B*T, d]
  bf16* p_out;              // [H, B, T, ld] or nullptr
  bf16* pd_out;             // [H, B, T, ld] or nullptr (no dropout)
  float drop_p;
  uint32_t thresh;
  unsigned long long seed;
  const unsigned long long* seed_ptr;
};

// ---- PTX wrappers (same conventions as gemm_tcgen05.cu) ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 28)) __trap();  // a protocol bug must surface as a launch failure, never as a hung GPU
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
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
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// SWIZZLE_128B shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D = f32, A = B = bf16, M = 128
__device__ __forceinline__ uint32_t make_idesc(int n, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTile >> 4) << 24);
}

// two adjacent 16-byte groups of a probability row: one 32-byte store when aligned, guarded per group against the row end
__device__ __forceinline__ void store_2x16(bf16* dst, const uint4 (&v)[2], int j, int ld) {
  if (j + 16 <= ld && ((reinterpret_cast<uintptr_t>(dst) & 31) == 0)) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(v[0].x), "r"(v[0].y), "r"(v[0].z),
                 "r"(v[0].w), "r"(v[1].x), "r"(v[1].y), "r"(v[1].z), "r"(v[1].w)
                 : "memory");
  } else {
    if (j < ld) *reinterpret_cast<uint4*>(dst) = v[0];
    if (j + 8 < ld) *reinterpret_cast<uint4*>(dst + 8) = v[1];
  }
}
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// Thread layout: warps 0-7 are softmax warps -- warp w owns TMEM lane quarter (w & 3) (rows 32 (w & 3) .. + 31 of the
// query tile) and column half (w >> 2) of the 128-key tile, so every query row is shared by two threads; warp 8 is the
// control warp (TMA + MMA issue + TMEM allocation).  The softmax warps release S / W as soon as they hold the logits in
// registers, so the tensor core works on tile it + 1 while they exponentiate tile it.
__global__ void __launch_bounds__(kThreads, 1)
attn_fused_fwd_kernel(const __grid_constant__ CUtensorMap tmQu, const __grid_constant__ CUtensorMap tmQv,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const __grid_constant__ CUtensorMap tmP, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar0 = sb + kOffBar;
  // mbarriers: Q | K+positions stage 0,1 | V stage 0,1 | S/W ready | S/W read (TMEM free) | A operand written | P V done
  const uint32_t barQ = bar0, barK0 = bar0 + 8, barV0 = bar0 + 24, barS = bar0 + 40, barT = bar0 + 48, barP = bar0 + 56,
                 barO = bar0 + 64;
  uint32_t* tmem_slot = (uint32_t*)(smem + kOffBar + 96);
  float* xch = (float*)(smem + kOffBar + 128);  // [2][128][2] (m, l) exchange between the two threads of a row

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T;
  const int i0 = qt * kTile;
  const int nkt = (T + kTile - 1) / kTile;  // key tiles
  const int total = 2 * nkt;                // pass 1 (statistics) + pass 2 (probabilities, P V)
  const int wbase = (T - 1) - i0 - (kTile - 1);  // relative position seen by (row 127, key 0); block a = wbase + 128 a

  esp_pdl_trigger();
  if (threadIdx.x == 256) {
    mbar_init(barQ, 1);
    mbar_init(barK0, 1);
    mbar_init(barK0 + 8, 1);
    mbar_init(barV0, 1);
    mbar_init(barV0 + 8, 1);
    mbar_init(barS, 1);
    mbar_init(barT, 8);  // one arrive per softmax warp
    mbar_init(barP, 8);
    mbar_init(barO, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQu) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  esp_pdl_wait();

  if (warp == 8) {
    // ============================== control: TMA loads + MMA issue (one thread) ==================================
    if (lane == 0) {
      const int pcol = h * p.pos_hstride;
      // K tile + relative-position blocks of iteration `it`.  Block a lives in ring slot a % 3; within a pass tile jt needs
      // blocks jt (loaded by the previous iteration) and jt + 1.
      auto load_k = [&](int it) {
        const int jt = it % nkt, st = it & 1;
        const bool first = jt == 0;
        const uint32_t bar = barK0 + 8 * st;
        mbar_expect_tx(bar, (uint32_t)kTileBytes * (first ? 3 : 2));
        tma_load_4d(sb + kOffK + st * kTileBytes, &tmK, bar, h * kHd, jt * kTile, b, 0);
        if (first) tma_load_4d(sb + kOffP + (jt % 3) * kTileBytes, &tmP, bar, pcol, wbase + jt * kTile, 0, 0);
        tma_load_4d(sb + kOffP + ((jt + 1) % 3) * kTileBytes, &tmP, bar, pcol, wbase + (jt + 1) * kTile, 0, 0);
      };
      const uint32_t id128 = make_idesc(128, false), id64 = make_idesc(kHd, true);
      auto issue_sw = [&](int it) {
        const int jt = it % nkt, st = it & 1;
        mbar_wait(barK0 + 8 * st, (it >> 1) & 1);
        if (it > 0) mbar_wait(barT, (it - 1) & 1);  // the softmax warps hold tile it-1's logits in registers
        tcgen05_fence_after();
        const uint32_t sk = sb + kOffK + st * kTileBytes;
        const uint32_t plo = sb + kOffP + (jt % 3) * kTileBytes, phi = sb + kOffP + ((jt + 1) % 3) * kTileBytes;
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_bf16(tmem + kColS, make_sdesc(sb + kOffQu + k * 32, 16, 1024), make_sdesc(sk + k * 32, 16, 1024), id128, k);
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_bf16(tmem + kColW, make_sdesc(sb + kOffQv + k * 32, 16, 1024), make_sdesc(plo + k * 32, 16, 1024), id128, k);
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_bf16(tmem + kColW + 128, make_sdesc(sb + kOffQv + k * 32, 16, 1024), make_sdesc(phi + k * 32, 16, 1024),
                    id128, k);
        tcgen05_commit(barS);
      };
      auto issue_pv = [&](int jt) {
        mbar_wait(barV0 + 8 * (jt & 1), (jt >> 1) & 1);
        mbar_wait(barP, jt & 1);  // the dropped probabilities of tile jt are in shared memory
        tcgen05_fence_after();
        const uint32_t sv = sb + kOffV + (jt & 1) * kTileBytes;
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k)
          umma_bf16(tmem + kColO, make_sdesc(sb + kOffA + (k >> 2) * kTileBytes + (k & 3) * 32, 16, 1024),
                    make_sdesc(sv + k * 2048, kTileBytes, 1024), id64, (jt > 0 || k > 0) ? 1u : 0u);
        tcgen05_commit(barO);
      };
      mbar_expect_tx(barQ, 2 * kTileBytes);
      tma_load_4d(sb + kOffQu, &tmQu, barQ, h * kHd, i0, b, 0);
      tma_load_4d(sb + kOffQv, &tmQv, barQ, h * kHd, i0, b, 0);
      load_k(0);
      mbar_wait(barQ, 0);
      for (int it = 0; it < total; ++it) {
        issue_sw(it);
        if (it + 1 < total) {
          // K stage (it+1)&1 and ring slot (jt+2)%3 were last read by the S/W MMAs of iteration it-1, which completed
          // before barT(it-1) could be observed.  Across the pass boundary the ring restarts at block 0 and would
          // overwrite blocks the MMAs of THIS iteration read: wait for them first.
          if ((it + 1) % nkt == 0) mbar_wait(barT, it & 1);
          load_k(it + 1);
        }
        if (it >= nkt) {
          // V tile of this pass-2 tile (consumed by its P V one iteration later); its stage was last read by P V(jt-2)
          const int jt = it - nkt;
          if (jt >= 2) mbar_wait(barO, (jt - 2) & 1);
          const uint32_t bar = barV0 + 8 * (jt & 1);
          mbar_expect_tx(bar, kTileBytes);
          tma_load_4d(sb + kOffV + (jt & 1) * kTileBytes, &tmV, bar, h * kHd, jt * kTile, b, 0);
        }
        if (it > nkt) issue_pv(it - 1 - nkt);
      }
      issue_pv(nkt - 1);
    }
  } else {
    // ============================== softmax warps ===============================================================
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;  // query row inside the tile = TMEM lane
    const int qi = i0 + r;
    const bool row_ok = qi < T;
    const int klen = p.lens ? min(p.lens[b], T) : T;
    int kmin = 0, kmax = T;
    if (p.key_lo) {  // chunk-streaming / limited-context attention: a contiguous key range per query row
      const int qr = min(qi, T - 1);
      kmin = p.key_lo[qr];
      kmax = p.key_hi[qr];
    }
    const unsigned long long seed = p.seed + (p.seed_ptr ? *p.seed_ptr : 0ull);
    const float dscale = p.drop_p > 0.f ? 65536.f / (65536.f - (float)p.thresh) : 1.f;
    const long prow = ((long)h * p.B + b) * T + qi;  // row of the probability tensors
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    uint8_t* wrow = smem + kOffW + r * kWRowBytes;
    const int c_lo = 96 - 32 * q;  // first W column this lane quarter stages (160 columns from there)
    float m_run = -INFINITY, l_run = 0.f, inv_l = 0.f;
    const float kLog2e = 1.4426950408889634f;
    constexpr int kHalf = kTile / 2;

    for (int it = 0; it < total; ++it) {
      const int jt = it % nkt;
      const bool pass2 = it >= nkt;
      const int j0 = jt * kTile + hf * kHalf;  // first key of this thread's half tile
      if (it == 0) TLF(0);
      mbar_wait(barS, it & 1);
      TLF(1 + 3 * it);
      tcgen05_fence_after();
      // ---- stage the row's slice of W in shared memory as bf16: half 0 stages chunks 0-2 (all it reads itself),
      //      half 1 stages chunks 3-4 and additionally needs chunk 2 from its partner warp ----
      const int cb = hf ? 3 : 0, ce = hf ? 5 : 3;
#pragma unroll 1
      for (int c = cb; c < ce; ++c) {
        uint32_t w[32];
        tmem_ld32(lane_base + (uint32_t)(kColW + c_lo + 32 * c), w);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(w[8 * q4 + 0]), __uint_as_float(w[8 * q4 + 1]));
          o.y = pack_bf16x2(__uint_as_float(w[8 * q4 + 2]), __uint_as_float(w[8 * q4 + 3]));
          o.z = pack_bf16x2(__uint_as_float(w[8 * q4 + 4]), __uint_as_float(w[8 * q4 + 5]));
          o.w = pack_bf16x2(__uint_as_float(w[8 * q4 + 6]), __uint_as_float(w[8 * q4 + 7]));
          *reinterpret_cast<uint4*>(wrow + c * 64 + q4 * 16) = o;
        }
      }
      if (hf == 0) named_arrive(1 + q, 64);
      else named_sync(1 + q, 64);
      // ---- logits of the half row: s[c] = S[r, c] + W[r, 127 - r + c], masked beyond the utterance's keys ----
      // the window starts (31 - lane) + 64 hf bf16 into the staged slice: odd offsets are re-aligned with a funnel shift
      const int woff = 31 - lane + hf * kHalf;
      const uint32_t* wwords = reinterpret_cast<const uint32_t*>(wrow) + (woff >> 1);
      const bool odd = (woff & 1) != 0;
      float s[kHalf];
      float tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t sr[32];
        tmem_ld32(lane_base + (uint32_t)(kColS + hf * kHalf + 32 * c), sr);
        uint32_t ww[17];
#pragma unroll
        for (int x = 0; x < 17; ++x) ww[x] = wwords[16 * c + x];
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 16; ++x) {
          const uint32_t wv = odd ? __funnelshift_r(ww[x], ww[x + 1], 16) : ww[x];
          float lo, hi;
          unpack_bf16x2(wv, lo, hi);
          const int j = j0 + 32 * c + 2 * x;
          float a0 = __uint_as_float(sr[2 * x]) + lo, a1 = __uint_as_float(sr[2 * x + 1]) + hi;
          if (p.key_lo) {
            // hidden keys: the reference ADDS bf16(-1e4) to its bf16 scores (transformer_layer.py:189-192,
            // multihead_attention.py:835-839) -- finite, so a row whose visible keys are all padding still
            // normalises over the hidden ones
            if (j < kmin || j >= kmax) a0 = __bfloat162float(__float2bfloat16_rn(a0 - 9984.f));
            if (j + 1 < kmin || j + 1 >= kmax) a1 = __bfloat162float(__float2bfloat16_rn(a1 - 9984.f));
          }
          a0 = (j < klen) ? a0 : -INFINITY;      // key padding: -inf (:841-859)
          a1 = (j + 1 < klen) ? a1 : -INFINITY;
          s[32 * c + 2 * x] = a0;
          s[32 * c + 2 * x + 1] = a1;
          tmax = fmaxf(tmax, fmaxf(a0, a1));
        }
      }
      // S / W (TMEM) and the staged slice are consumed: the tensor core may start the next tile
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(barT);
      TLF(2 + 3 * it);
      if (!pass2) {
        // ---- pass 1: online maximum and normaliser of this thread's half of the row ----
        const float m_new = fmaxf(m_run, tmax);
        if (m_new > -INFINITY) {
          float acc = 0.f;
          const float mb = m_new * kLog2e;
#pragma unroll
          for (int c = 0; c < kHalf; ++c) acc += exp2f(fmaf(s[c], kLog2e, -mb));  // exp2(-inf) = 0 for masked keys
          l_run = l_run * exp2f((m_run - m_new) * kLog2e) + acc;
          m_run = m_new;
        }
        if (it == nkt - 1) {
          // merge the two halves of the row
          xch[(hf * kTile + r) * 2] = m_run;
          xch[(hf * kTile + r) * 2 + 1] = l_run;
          named_sync(5 + q, 64);
          const float m_o = xch[((hf ^ 1) * kTile + r) * 2], l_o = xch[((hf ^ 1) * kTile + r) * 2 + 1];
          const float m_all = fmaxf(m_run, m_o);
          float l_all = 0.f;
          if (m_run > -INFINITY) l_all += l_run * exp2f((m_run - m_all) * kLog2e);
          if (m_o > -INFINITY) l_all += l_o * exp2f((m_o - m_all) * kLog2e);
          m_run = m_all;
          inv_l = l_all > 0.f ? 1.f / l_all : 0.f;
        }
      } else {
        // ---- pass 2: normalised probabilities, dropout, A operand of P V ----
        const float mb = m_run * kLog2e;
        bf16* prp = p.p_out ? p.p_out + prow * p.ld + j0 : nullptr;
        bf16* pdp = p.pd_out ? p.pd_out + prow * p.ld + j0 : nullptr;
        uint8_t* arow = smem + kOffA + hf * kTileBytes + r * 128;  // 64-column chunk hf of the K-major A operand
        const int tj = it - nkt;
        if (tj > 0) mbar_wait(barO, (tj - 1) & 1);  // P V of the previous tile has finished reading the A operand
#pragma unroll
        for (int g2 = 0; g2 < kHalf / 16; ++g2) {
          // 16 keys per iteration so that every global store is one full 32-byte sector (sub-sector writes are slow)
          uint4 pr4[2], pd4[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int g = 2 * g2 + u;
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float pv = exp2f(fmaf(s[8 * g + e], kLog2e, -mb)) * inv_l;
              o[e] = bf2f(f2bf(pv));  // softmax in fp32, cast to the model dtype (fairseq/utils.py:514-525 + .type_as)
            }
            pr4[u].x = pack_bf16x2(o[0], o[1]); pr4[u].y = pack_bf16x2(o[2], o[3]);
            pr4[u].z = pack_bf16x2(o[4], o[5]); pr4[u].w = pack_bf16x2(o[6], o[7]);
            if (p.drop_p > 0.f) {
              bool keep[8];
              esp_keep8(seed, (unsigned long long)prow * p.ld + j0 + 8 * g, p.thresh, keep);
#pragma unroll
              for (int e = 0; e < 8; ++e) o[e] = keep[e] ? o[e] * dscale : 0.f;
            }
            pd4[u].x = pack_bf16x2(o[0], o[1]); pd4[u].y = pack_bf16x2(o[2], o[3]);
            pd4[u].z = pack_bf16x2(o[4], o[5]); pd4[u].w = pack_bf16x2(o[6], o[7]);
            // K-major SWIZZLE_128B: 16-byte unit g XOR (row & 7) inside the row's 128 bytes
            *reinterpret_cast<uint4*>(arow + ((g ^ (r & 7)) << 4)) = pd4[u];
          }
          const int j = j0 + 16 * g2;
          if (row_ok) {
            if (prp) store_2x16(prp + 16 * g2, pr4, j, p.ld);
            if (pdp) store_2x16(pdp + 16 * g2, pd4, j, p.ld);
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
        __syncwarp();
        if (lane == 0) mbar_arrive(barP);
      }
      TLF(3 + 3 * it);
    }
    // ---- epilogue: O (already normalised) -> ctx; this thread stores 32 of the row's 64 values ----
    mbar_wait(barO, (nkt - 1) & 1);
    tcgen05_fence_after();
    {
      uint32_t o[32];
      tmem_ld32(lane_base + (uint32_t)(kColO + 32 * hf), o);
      tmem_ld_wait();
      if (row_ok) {
        bf16* dst = p.ctx + ((long)b * T + qi) * p.d + h * kHd + 32 * hf;
#pragma unroll
        for (int q2 = 0; q2 < 2; ++q2) {
          uint4 v4[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int o0 = 16 * q2 + 8 * u;
            v4[u].x = pack_bf16x2(__uint_as_float(o[o0 + 0]), __uint_as_float(o[o0 + 1]));
            v4[u].y = pack_bf16x2(__uint_as_float(o[o0 + 2]), __uint_as_float(o[o0 + 3]));
            v4[u].z = pack_bf16x2(__uint_as_float(o[o0 + 4]), __uint_as_float(o[o0 + 5]));
            v4[u].w = pack_bf16x2(__uint_as_float(o[o0 + 6]), __uint_as_float(o[o0 + 7]));
          }
          store_2x16(dst + 16 * q2, v4, 0, 1 << 30);
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 8) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
  }
}

}  // namespace

int esp_make_tmap_bf16(CUtensorMap* tm, const void* base, long inner, long rows, long ld, int nb1, long s1, int nb2,
                       long s2, int box_rows);  // gemm_tcgen05.cu

extern "C" int esp_attn_fused_fwd(const void* qu, const void* qv, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                                  const void* pos, int64_t ldpos, int32_t pos_hstride, int32_t B, int32_t T, int32_t H,
                                  int32_t head_dim, const int32_t* lens, const int32_t* key_lo, const int32_t* key_hi,
                                  void* ctx, int64_t ldctx, void* p_out, void* pd_out, int32_t ldp, float drop_p,
                                  uint64_t seed, const uint64_t* seed_ptr, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ESP_CHECK(head_dim == kHd, "fused attention is built for head_dim 64 (got %d)", head_dim);
  ESP_CHECK(B >= 0 && T >= 0 && H > 0, "bad attention shape");
  if (B == 0 || T == 0) return 0;
  ESP_CHECK(qu && qv && k && v && pos && ctx, "null pointer passed to esp_attn_fused_fwd");
  ESP_CHECK(p_out == nullptr || (ldp >= T && ldp % 8 == 0), "probability row stride must be a multiple of 8 and >= T");
  ESP_CHECK(drop_p <= 0.f || p_out == nullptr || pd_out != nullptr, "dropout with saved probabilities needs pd_out");
  ESP_CHECK(ldctx == (int64_t)H * kHd, "ctx must be [B*T, H*64] contiguous");
  ESP_CHECK(B <= 65535 && H <= 65535, "grid limits");
  CUtensorMap tqu, tqv, tk, tv, tp;
  int rc;
  if ((rc = esp_make_tmap_bf16(&tqu, qu, (long)H * kHd, T, ldq, B, (long)T * ldq, 1, 0, kTile))) return rc;
  if ((rc = esp_make_tmap_bf16(&tqv, qv, (long)H * kHd, T, ldq, B, (long)T * ldq, 1, 0, kTile))) return rc;
  if ((rc = esp_make_tmap_bf16(&tk, k, (long)H * kHd, T, ldkv, B, (long)T * ldkv, 1, 0, kTile))) return rc;
  if ((rc = esp_make_tmap_bf16(&tv, v, (long)H * kHd, T, ldkv, B, (long)T * ldkv, 1, 0, kTile))) return rc;
  const long pos_cols = pos_hstride ? (long)H * kHd : kHd;
  if ((rc = esp_make_tmap_bf16(&tp, pos, pos_cols, 2L * T - 1, ldpos, 1, 0, 1, 0, kTile))) return rc;
  Params pr;
  {
    const char* tl = getenv("ESP_ATTN_FWD_TIMELINE");
    pr.timeline = (tl && tl[0] == '1') ? 1 : 0;
  }
  pr.B = B; pr.T = T; pr.H = H; pr.d = (int)ldctx; pr.ld = ldp; pr.pos_hstride = pos_hstride;
  ESP_CHECK((key_lo == nullptr) == (key_hi == nullptr), "key_lo and key_hi must be given together");
  pr.key_lo = key_lo; pr.key_hi = key_hi;
  pr.lens = lens; pr.ctx = (bf16*)ctx; pr.p_out = (bf16*)p_out; pr.pd_out = drop_p > 0.f ? (bf16*)pd_out : nullptr;
  pr.drop_p = drop_p; pr.thresh = esp_dropout_thresh(drop_p); pr.seed = seed;
  pr.seed_ptr = (const unsigned long long*)seed_ptr;
  static bool configured = false;
  if (!configured) {
    ESP_CUDA(cudaFuncSetAttribute(attn_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  dim3 grid((T + kTile - 1) / kTile, H, B);
  esp_launch(attn_fused_fwd_kernel, grid, kThreads, kSmemBytes, st, tqu, tqv, tk, tv, tp, pr);
  ESP_LAUNCH_CHECK();
  esp_count_launch(1);
  return 0;
}

// debugging aid: the %globaltimer stamps (ns) of CTA (0, 0, 0) of the last esp_attn_fused_fwd launch with
// ESP_ATTN_FWD_TIMELINE=1 (profiles/attn_fwd_timeline.py)
extern "C" int esp_attn_fwd_timeline(unsigned long long* out128) {
  ESP_CUDA(cudaDeviceSynchronize());
  ESP_CUDA(cudaMemcpyFromSymbol(out128, g_fwd_timeline, sizeof(unsigned long long) * 128));
  return 0;
}
