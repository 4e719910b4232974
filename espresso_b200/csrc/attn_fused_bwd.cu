// espresso_b200 -- fused relative-position self-attention BACKWARD (score side) for sm_100a.
//
// Backward of  ctx = dropout(softmax(qu k^T + skew(qv pos^T))) v   (fairseq/modules/multihead_attention.py:788-897) given
// dctx and the probabilities saved by esp_attn_fused_fwd.  One CTA per (128-key tile, head, utterance) walks the query
// tiles and keeps everything between the saved probabilities and the four results on chip:
//
//   dPd = dctx_i v_j^T                      tcgen05.mma into TMEM (128 x 128 fp32, double-buffered)
//   dS  = P . (mask/keep . dPd - D_i)       softmax warps, registers;  D_i = sum_e dctx[i,e] ctx[i,e]  (= sum_j P dP)
//   dV_j += Pd^T dctx_i ,  dK_j += dS^T qu_i   tcgen05.mma, accumulators stay in TMEM over the whole query loop
//
// and writes dS [H,B,T,ld] and its skewed copy dBD[., i, (T-1)-i+j] (zeros elsewhere; inverse of the Transformer-XL
// shift, :824-830) once, for the three remaining GEMMs (dq_u = dS k, dq_v = dBD pos, dpos = dBD^T q_v).
// It replaces four launches of the unfused chain (the dPd GEMM, esp_attn_softmax_bwd, the dV and dK GEMMs) and the HBM
// round trips of dPd (write + read), P_drop and two reads of dS.
//
// Operand tiles are TMA boxes (SWIZZLE_128B) of 64 columns x 128 rows.  The probability tile [128 queries x 128 keys]
// (two 64-key chunks) is at the same time
//   * the MN-major A operand of an M = keys, K = queries product (dV, dK: the transposes are free), and
//   * the place where the softmax warps overwrite P with dS (same rows, same 16-byte units).
#include "common.cuh"
#include "espresso_b200.h"
#include <cuda.h>
#include <stdlib.h>

void esp_count_launch(int n);

namespace {

constexpr int kTile = 128;
constexpr int kHd = 64;
constexpr int kThreads = 288;  // warps 0-7: dS warps (TMEM lane quarter x key half), warp 8: TMA + MMA issue + TMEM alloc
constexpr int kTileBytes = kTile * kHd * 2;  // 16 KB: one 64-column x 128-row box
constexpr int kOffV = 0;
constexpr int kOffStage = kTileBytes;
constexpr int kStDO = 0, kStQu = kTileBytes, kStP = 2 * kTileBytes, kStX = 4 * kTileBytes;
constexpr int kStageBytes = 6 * kTileBytes;  // dctx_i, qu_i, P (2 chunks), P_drop (2 chunks)
constexpr int kOffBar = kOffStage + 2 * kStageBytes;
constexpr int kSmemBytes = kOffBar + 256 + 1024;
constexpr int kColDP = 0, kColDV = 256, kColDK = 320, kTmemCols = 512;

// optional timeline of CTA (0, 0, 0) (ESP_ATTN_BWD_TIMELINE=1; esp_attn_bwd_timeline reads it): %globaltimer stamps of the
// control thread (slots 0-63) and of dS warp 0 (slots 64-127)
__device__ unsigned long long g_timeline[128];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TL(slot)                                                                                  \
  do {                                                                                            \
    if ((p.timeline & 1) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_timeline[(slot)] = gtimer(); \
  } while (0)

struct Params {
  int timeline;
  int B, T, H, ld, ldp;
  const float* D;            // [H, B, T] row dots
  bf16* ds_out;              // [H, B, T, ld]
  bf16* dbd_out;             // [H, B, T, ldp]
  bf16* dk_out;              // [B*T, ld_out] (head h at column h*64)
  bf16* dv_out;
  long ld_out;
  float drop_p;
  uint32_t thresh;
  unsigned long long seed;
  const unsigned long long* seed_ptr;
};

// ---- PTX wrappers (same conventions as attn_fused.cu / gemm_tcgen05.cu) ----------------------------------------------
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
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D = f32, A = B = bf16, M = 128; operand majors as flags (0 = K-major, 1 = MN-major)
__device__ __forceinline__ uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
}
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void store_2x16(bf16* dst, const uint4& a, const uint4& b, int j, int ld) {
  if (j + 16 <= ld && ((reinterpret_cast<uintptr_t>(dst) & 31) == 0)) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
  } else {
    if (j < ld) *reinterpret_cast<uint4*>(dst) = a;
    if (j + 8 < ld) *reinterpret_cast<uint4*>(dst + 8) = b;
  }
}

// D[h, b, t] = sum_e dctx[b*T+t, h*64+e] * ctx[b*T+t, h*64+e]: one warp per row of [B*T, H*64], 8 lanes per head
__global__ void __launch_bounds__(256)
attn_rowdot_kernel(const bf16* __restrict__ a, const bf16* __restrict__ c, long R, int T, int H, long ld,
                   float* __restrict__ D) {
  esp_pdl();
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R) return;
  const int b = (int)(row / T), t = (int)(row % T), B = (int)(R / T);
  for (int h0 = 0; h0 < H; h0 += 4) {  // 32 lanes x 8 elements = 4 heads per trip
    const int h = h0 + (lane >> 3);
    float s = 0.f;
    if (h < H) {
      const uint4 qa = *reinterpret_cast<const uint4*>(a + row * ld + h0 * kHd + lane * 8);
      const uint4 qc = *reinterpret_cast<const uint4*>(c + row * ld + h0 * kHd + lane * 8);
      float x0, x1, y0, y1;
      unpack_bf16x2(qa.x, x0, x1); unpack_bf16x2(qc.x, y0, y1); s += x0 * y0 + x1 * y1;
      unpack_bf16x2(qa.y, x0, x1); unpack_bf16x2(qc.y, y0, y1); s += x0 * y0 + x1 * y1;
      unpack_bf16x2(qa.z, x0, x1); unpack_bf16x2(qc.z, y0, y1); s += x0 * y0 + x1 * y1;
      unpack_bf16x2(qa.w, x0, x1); unpack_bf16x2(qc.w, y0, y1); s += x0 * y0 + x1 * y1;
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if ((lane & 7) == 0 && h < H) D[((long)h * B + b) * T + t] = s;
  }
}

// dBD[row, (T-1) - i + j] = dS[row, j], zeros elsewhere (inverse of the Transformer-XL shift, :824-830): one warp per row
// (h, b, i).  The row is staged in shared memory with 16-byte loads; every destination-aligned group of 8 columns is then the
// tail of one source unit and the head of the next (the split point eo = ((T-1)-i) mod 8 is the same for the whole row), so a
// lane builds a group from two 16-byte shared-memory loads and a funnel shift and writes it with one aligned 16-byte store.
// Plenty of resident warps hide the latencies -- which the fused kernel above, with its eight dS warps per SM, cannot.
constexpr int kSkewWarps = 8;
constexpr int kSkewMaxUnits = 130;  // T <= 1024: 128 source units + one zero unit on either side
__global__ void __launch_bounds__(kSkewWarps * 32)
attn_skew_kernel(const bf16* __restrict__ ds, bf16* __restrict__ dbd, long rows, int T, int ld, int ldp) {
  esp_pdl();
  __shared__ uint4 stage[kSkewWarps][kSkewMaxUnits];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long row = (long)blockIdx.x * kSkewWarps + w;
  if (row >= rows) return;
  const int i = (int)(row % T);
  const int nu = (T + 7) >> 3;  // source units that hold keys (ld >= 8 nu; columns T..8nu-1 of dS are zero)
  uint4* st = stage[w];
  const uint4* src = reinterpret_cast<const uint4*>(ds + row * (long)ld);
  // staged units: index 0 = zeros, 1..nu = the row, nu+1 = zeros
  for (int u = lane; u < nu + 2; u += 32) st[u] = (u >= 1 && u <= nu) ? src[u - 1] : make_uint4(0, 0, 0, 0);
  __syncwarp();
  const int cs = (T - 1) - i;      // destination column of key 0
  const int a = cs & ~7, eo = cs & 7;
  const int ga = a >> 3;           // destination group that holds key 0
  uint4* dst = reinterpret_cast<uint4*>(dbd + row * (long)ldp);
  const int ngroups = ldp >> 3;
  for (int gd = lane; gd < ngroups; gd += 32) {
    const int g = gd - ga;  // group index inside the window: source units g - 1 (tail) and g (head)
    uint4 o = make_uint4(0, 0, 0, 0);
    if (g >= 0 && g <= nu) {
      const uint4 u0 = st[g], u1 = st[g + 1];  // staged index = unit + 1
      const uint32_t x[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
      switch (8 - eo) {  // first element of the group inside the 16-element pair (warp-uniform)
        case 8: o = u1; break;
        case 7: o = make_uint4(__funnelshift_r(x[3], x[4], 16), __funnelshift_r(x[4], x[5], 16), __funnelshift_r(x[5], x[6], 16), __funnelshift_r(x[6], x[7], 16)); break;
        case 6: o = make_uint4(x[3], x[4], x[5], x[6]); break;
        case 5: o = make_uint4(__funnelshift_r(x[2], x[3], 16), __funnelshift_r(x[3], x[4], 16), __funnelshift_r(x[4], x[5], 16), __funnelshift_r(x[5], x[6], 16)); break;
        case 4: o = make_uint4(x[2], x[3], x[4], x[5]); break;
        case 3: o = make_uint4(__funnelshift_r(x[1], x[2], 16), __funnelshift_r(x[2], x[3], 16), __funnelshift_r(x[3], x[4], 16), __funnelshift_r(x[4], x[5], 16)); break;
        case 2: o = make_uint4(x[1], x[2], x[3], x[4]); break;
        default: o = make_uint4(__funnelshift_r(x[0], x[1], 16), __funnelshift_r(x[1], x[2], 16), __funnelshift_r(x[2], x[3], 16), __funnelshift_r(x[3], x[4], 16)); break;
      }
    }
    dst[gd] = o;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
attn_fused_bwd_kernel(const __grid_constant__ CUtensorMap tmDO, const __grid_constant__ CUtensorMap tmQu,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmP,
                      const __grid_constant__ CUtensorMap tmPd, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t sb = smem_u32(smem);
  const uint32_t bar0 = sb + kOffBar;
  // mbarriers: V | stage loaded x2 | dPd ready x2 | dS written x2 | stage consumed by the MMAs x2 | stage read by the
  // skewed stores x2 | accumulators final
  const uint32_t barV = bar0, barL = bar0 + 8, barDP = bar0 + 24, barDS = bar0 + 40, barFree = bar0 + 56, barRD = bar0 + 72,
                 barAcc = bar0 + 88;
  uint32_t* tmem_slot = (uint32_t*)(smem + kOffBar + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int T = p.T;
  const int j0 = jt * kTile;
  const int nqt = (T + kTile - 1) / kTile;  // query tiles
  const int nkt = gridDim.x;

  esp_pdl_trigger();
  if (threadIdx.x == 256) {
    mbar_init(barV, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(barL + 8 * s, 1);
      mbar_init(barDP + 8 * s, 1);
      mbar_init(barDS + 8 * s, 8);  // one arrive per dS warp
      mbar_init(barFree + 8 * s, 1);
      mbar_init(barRD + 8 * s, 8);
    }
    mbar_init(barAcc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmDO) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmPd) : "memory");
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
  if (threadIdx.x == 256) TL(0);
  esp_pdl_wait();
  if (threadIdx.x == 256) TL(1);

  if (warp == 8) {
    // ============================== control: TMA loads + MMA issue (one thread) ==================================
    if (lane == 0) {
      auto load_stage = [&](int i) {
        const int s = i & 1;
        const uint32_t bar = barL + 8 * s, st = sb + kOffStage + s * kStageBytes;
        mbar_expect_tx(bar, (uint32_t)kStageBytes);
        tma_load_4d(st + kStDO, &tmDO, bar, h * kHd, i * kTile, b, 0);
        tma_load_4d(st + kStQu, &tmQu, bar, h * kHd, i * kTile, b, 0);
        tma_load_4d(st + kStP, &tmP, bar, j0, i * kTile, b, h);
        tma_load_4d(st + kStP + kTileBytes, &tmP, bar, j0 + 64, i * kTile, b, h);
        tma_load_4d(st + kStX, &tmPd, bar, j0, i * kTile, b, h);
        tma_load_4d(st + kStX + kTileBytes, &tmPd, bar, j0 + 64, i * kTile, b, h);
      };
      const uint32_t id_dp = make_idesc(128, false, false), id_acc = make_idesc(kHd, true, true);
      // dV_j += Pd^T dctx_i ; dK_j += dS^T qu_i : M = keys (two 64-key chunks, 16 KB apart), K = the tile's 128 queries
      auto issue_acc = [&](int i) {
        const int s = i & 1;
        const uint32_t st = sb + kOffStage + s * kStageBytes;
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k)
          umma_bf16(tmem + kColDV, make_sdesc(st + kStX + k * 2048, kTileBytes, 1024),
                    make_sdesc(st + kStDO + k * 2048, kTileBytes, 1024), id_acc, (i > 0 || k > 0) ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < kTile / 16; ++k)
          umma_bf16(tmem + kColDK, make_sdesc(st + kStP + k * 2048, kTileBytes, 1024),
                    make_sdesc(st + kStQu + k * 2048, kTileBytes, 1024), id_acc, (i > 0 || k > 0) ? 1u : 0u);
      };
      mbar_expect_tx(barV, kTileBytes);
      tma_load_4d(sb + kOffV, &tmV, barV, h * kHd, j0, b, 0);
      load_stage(0);
      if (nqt > 1) load_stage(1);
      mbar_wait(barV, 0);
      TL(2);
      for (int i = 0; i < nqt; ++i) {
        const int s = i & 1;
        mbar_wait(barL + 8 * s, (i >> 1) & 1);
        if (i < 8) TL(4 + 4 * i);
        tcgen05_fence_after();
        // dPd = dctx_i v_j^T into TMEM buffer s (its previous contents, tile i-2, were read before barDS(i-2))
        const uint32_t st = sb + kOffStage + s * kStageBytes;
#pragma unroll
        for (int k = 0; k < kHd / 16; ++k)
          umma_bf16(tmem + kColDP + s * kTile, make_sdesc(st + kStDO + k * 32, 16, 1024),
                    make_sdesc(sb + kOffV + k * 32, 16, 1024), id_dp, k);
        tcgen05_commit(barDP + 8 * s);
        if (i >= 1) {
          const int sp = (i - 1) & 1;
          const uint32_t ph = ((i - 1) >> 1) & 1;
          mbar_wait(barDS + 8 * sp, ph);  // dS of tile i-1 is in shared memory
          if (i < 8) TL(5 + 4 * i);
          tcgen05_fence_after();
          issue_acc(i - 1);
          tcgen05_commit(barFree + 8 * sp);
          if (i + 1 < nqt) {
            mbar_wait(barFree + 8 * sp, ph);  // the MMAs have read the stage ...
            if (i < 8) TL(6 + 4 * i);
            mbar_wait(barRD + 8 * sp, ph);    // ... and so have the skewed stores
            if (i < 8) TL(7 + 4 * i);
            load_stage(i + 1);
          }
        }
      }
      {
        const int sp = (nqt - 1) & 1;
        mbar_wait(barDS + 8 * sp, ((nqt - 1) >> 1) & 1);
        tcgen05_fence_after();
        issue_acc(nqt - 1);
        tcgen05_commit(barAcc);
      }
    }
  } else {
    // ============================== dS warps =====================================================================
    const int q = warp & 3, hf = warp >> 2;
    const int r = q * 32 + lane;  // query row inside the tile = TMEM lane
    const unsigned long long seed = p.seed + (p.seed_ptr ? *p.seed_ptr : 0ull);
    const float dscale = p.drop_p > 0.f ? 65536.f / (65536.f - (float)p.thresh) : 1.f;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    const int jh = j0 + hf * 64;  // first key of this thread's half row
    const long hb = (long)h * p.B + b;

    for (int i = 0; i < nqt; ++i) {
      const int s = i & 1;
      const int qi = i * kTile + r;
      const bool row_ok = qi < T;
      const long prow = hb * T + qi;
      const float Dv = row_ok ? p.D[prow] : 0.f;
      uint8_t* stg = smem + kOffStage + s * kStageBytes;
      uint8_t* prs = stg + kStP + hf * kTileBytes + r * 128;  // this thread's 64 keys of row r: 8 swizzled 16-byte units
      mbar_wait(barDP + 8 * s, (i >> 1) & 1);
      if (threadIdx.x == 0 && i < 8) TL(64 + 4 * i);
      tcgen05_fence_after();
      uint4 ds4[8];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t dp[32];
        tmem_ld32(lane_base + (uint32_t)(kColDP + s * kTile + hf * 64 + 32 * c), dp);
        uint4 pu[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) pu[u] = *reinterpret_cast<const uint4*>(prs + (((4 * c + u) ^ (r & 7)) << 4));
        tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int g = 4 * c + u;
          float pv[8];
          unpack_bf16x2(pu[u].x, pv[0], pv[1]); unpack_bf16x2(pu[u].y, pv[2], pv[3]);
          unpack_bf16x2(pu[u].z, pv[4], pv[5]); unpack_bf16x2(pu[u].w, pv[6], pv[7]);
          bool keep[8];
          if (p.drop_p > 0.f) esp_keep8(seed, (unsigned long long)prow * p.ld + jh + 8 * g, p.thresh, keep);
          float o[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float d = __uint_as_float(dp[8 * u + e]);
            if (p.drop_p > 0.f) d = keep[e] ? d * dscale : 0.f;
            o[e] = pv[e] * (d - Dv);
          }
          ds4[g].x = pack_bf16x2(o[0], o[1]); ds4[g].y = pack_bf16x2(o[2], o[3]);
          ds4[g].z = pack_bf16x2(o[4], o[5]); ds4[g].w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(prs + ((g ^ (r & 7)) << 4)) = ds4[g];  // dS replaces P in the operand tile
        }
      }
      // the accumulating MMAs may start as soon as all eight warps have written their part
      tcgen05_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(barDS + 8 * s);
      if (threadIdx.x == 0 && i < 8) TL(65 + 4 * i);
      // ---- dS rows (un-skewed) straight from registers: 64 keys = 128 contiguous bytes ----
      if (row_ok && !(p.timeline & 2)) {
        bf16* dsp = p.ds_out + prow * p.ld + jh;
#pragma unroll
        for (int g2 = 0; g2 < 4; ++g2) store_2x16(dsp + 16 * g2, ds4[2 * g2], ds4[2 * g2 + 1], jh + 16 * g2, p.ld);
      }
      // ---- skewed copy: dBD[row, (T-1) - qi + j] = dS[row, j]; the first key tile also writes the zeros in front of the
      //      row's window, the last one those behind it.  Both halves of a row must be in shared memory first. ----
      named_sync(1 + q, 64);
      if (threadIdx.x == 0 && i < 8) TL(66 + 4 * i);
      {
        // This warp stores rows q*32 + hf*16 + t, t = 0..15, of the tile.  Four rows per trip, everything per-row derived
        // from the trip's base with compile-time offsets, so that the four load -> funnel -> store chains interleave (a
        // warp issues dependent instructions several cycles apart and only eight warps share the SM).
        const int rbase = q * 32 + hf * 16;
        const int qr0 = i * kTile + rbase;
        const uint32_t tile_u32 = sb + kOffStage + s * kStageBytes + kStP + rbase * 128;
        bf16* const drow0 = p.dbd_out + (hb * T + qr0) * (long)p.ldp;
        const int cs0 = (T - 1) - qr0 + j0;  // destination column of key j0 for row t = 0; one less per row
        const bool first = jt == 0, last = jt == nkt - 1;
        const int g = lane;                  // destination-aligned group of 8 columns inside the row's window (0..16)
        if (p.dbd_out != nullptr && !(p.timeline & 8)) {
#pragma unroll 1
        for (int t0 = 0; t0 < 16; t0 += 4) {
          uint4 u0[4], u1[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int t = t0 + x;
            const uint32_t srow = tile_u32 + t * 128;
            const int sw = t & 7;  // (rbase + t) & 7: rbase is a multiple of 8
            u0[x] = make_uint4(0, 0, 0, 0);
            u1[x] = make_uint4(0, 0, 0, 0);
            // Destination-aligned groups of 8 columns (16 bytes): group g starts at a + 8 g, a = cs rounded down to 8; its
            // values are source elements 8 g - eo .. 8 g - eo + 7 (eo = cs - a): the tail of source unit g - 1 and the head
            // of unit g -- two 16-byte shared-memory loads and a funnel shift.
            if (g >= 1 && g <= 16)
              asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u0[x].x), "=r"(u0[x].y), "=r"(u0[x].z), "=r"(u0[x].w)
                  : "r"(srow + (uint32_t)(((g - 1) >> 3) * kTileBytes + ((((g - 1) & 7) ^ sw) << 4))));
            if (g <= 15)
              asm("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u1[x].x), "=r"(u1[x].y), "=r"(u1[x].z), "=r"(u1[x].w)
                  : "r"(srow + (uint32_t)((g >> 3) * kTileBytes + (((g & 7) ^ sw) << 4))));
          }
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            const int t = t0 + x;
            if (qr0 + t >= T) continue;
            bf16* drow = drow0 + (long)t * p.ldp;
            const int cs = cs0 - t;
            const int lo = first ? 0 : cs;
            const int hi = last ? p.ldp : cs + kTile;
            const int a = cs & ~7, eo = cs & 7;
            if (g <= 16) {
              const int c0 = a + 8 * g;
              const uint32_t w[8] = {u0[x].x, u0[x].y, u0[x].z, u0[x].w, u1[x].x, u1[x].y, u1[x].z, u1[x].w};
              uint4 o;
              switch (8 - eo) {  // first element of the group inside the 16-element pair (warp-uniform)
                case 8: o = u1[x]; break;
                case 7: o = make_uint4(__funnelshift_r(w[3], w[4], 16), __funnelshift_r(w[4], w[5], 16), __funnelshift_r(w[5], w[6], 16), __funnelshift_r(w[6], w[7], 16)); break;
                case 6: o = make_uint4(w[3], w[4], w[5], w[6]); break;
                case 5: o = make_uint4(__funnelshift_r(w[2], w[3], 16), __funnelshift_r(w[3], w[4], 16), __funnelshift_r(w[4], w[5], 16), __funnelshift_r(w[5], w[6], 16)); break;
                case 4: o = make_uint4(w[2], w[3], w[4], w[5]); break;
                case 3: o = make_uint4(__funnelshift_r(w[1], w[2], 16), __funnelshift_r(w[2], w[3], 16), __funnelshift_r(w[3], w[4], 16), __funnelshift_r(w[4], w[5], 16)); break;
                case 2: o = make_uint4(w[1], w[2], w[3], w[4]); break;
                default: o = make_uint4(__funnelshift_r(w[0], w[1], 16), __funnelshift_r(w[1], w[2], 16), __funnelshift_r(w[2], w[3], 16), __funnelshift_r(w[3], w[4], 16)); break;
              }
              if (p.timeline & 4) {
                if (o.x == 0x12345678u) drow[0] = f2bf(0.f);  // experiment: keep the loads alive, no stores
              } else if (c0 >= lo && c0 + 8 <= hi) {
                *reinterpret_cast<uint4*>(drow + c0) = o;
              } else {  // first / last group of the window: only the columns that belong to this tile
                const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int e = 0; e < 8; ++e)
                  if (c0 + e >= lo && c0 + e < hi)
                    reinterpret_cast<unsigned short*>(drow)[c0 + e] = (unsigned short)(ow[e >> 1] >> (16 * (e & 1)));
              }
            }
            // zeros in front of the first tile's window and behind the last tile's (whole 16-byte groups; the groups that
            // straddle the window were completed above)
            if (first)
              for (int c0 = 8 * lane; c0 + 8 <= a; c0 += 256) *reinterpret_cast<uint4*>(drow + c0) = make_uint4(0, 0, 0, 0);
            if (last)
              for (int c0 = a + 8 * 17 + 8 * lane; c0 < hi; c0 += 256) *reinterpret_cast<uint4*>(drow + c0) = make_uint4(0, 0, 0, 0);
          }
        }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(barRD + 8 * s);
      if (threadIdx.x == 0 && i < 8) TL(67 + 4 * i);
    }
    // ---- epilogue: dV_j, dK_j (fp32 in TMEM) -> bf16 rows of the fused dqkv buffer; this thread stores 32 of 64 dims ----
    mbar_wait(barAcc, 0);
    if (threadIdx.x == 0) TL(100);
    tcgen05_fence_after();
    {
      const int kj = j0 + r;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        uint32_t o[32];
        tmem_ld32(lane_base + (uint32_t)((w == 0 ? kColDV : kColDK) + 32 * hf), o);
        tmem_ld_wait();
        if (kj < T) {
          bf16* dst = (w == 0 ? p.dv_out : p.dk_out) + ((long)b * T + kj) * p.ld_out + h * kHd + 32 * hf;
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
            store_2x16(dst + 16 * q2, v4[0], v4[1], 0, 1 << 30);
          }
        }
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TL(101);
  if (warp == 8) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
  }
}

}  // namespace

// debugging aid: the %globaltimer stamps (ns) of CTA (0, 0, 0) of the last esp_attn_fused_bwd launch with
// ESP_ATTN_BWD_TIMELINE=1 (profiles/attn_bwd_timeline.py)
extern "C" int esp_attn_bwd_timeline(unsigned long long* out128) {
  ESP_CUDA(cudaDeviceSynchronize());
  ESP_CUDA(cudaMemcpyFromSymbol(out128, g_timeline, sizeof(unsigned long long) * 128));
  return 0;
}

int esp_make_tmap_bf16(CUtensorMap* tm, const void* base, long inner, long rows, long ld, int nb1, long s1, int nb2,
                       long s2, int box_rows);  // gemm_tcgen05.cu

extern "C" int esp_attn_fused_bwd(const void* dctx, const void* ctx, int64_t ldctx, const void* qu, int64_t ldq, const void* v,
                                  int64_t ldkv, const void* p, const void* pd, int32_t ldp_probs, int32_t B, int32_t T,
                                  int32_t H, int32_t head_dim, float drop_p, uint64_t seed, const uint64_t* seed_ptr,
                                  float* rowdot_ws, void* ds, void* dbd, int32_t ldbd, void* dk, void* dv, int64_t ld_out,
                                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ESP_CHECK(head_dim == kHd, "fused attention backward is built for head_dim 64 (got %d)", head_dim);
  ESP_CHECK(B >= 0 && T >= 0 && H > 0, "bad attention shape");
  if (B == 0 || T == 0) return 0;
  ESP_CHECK(dctx && ctx && qu && v && p && pd && rowdot_ws && ds && dbd && dk && dv, "null pointer passed to esp_attn_fused_bwd");
  ESP_CHECK(ldp_probs >= T && ldp_probs % 8 == 0, "probability row stride must be a multiple of 8 and >= T");
  ESP_CHECK(ldbd >= 2 * T - 1 && ldbd % 8 == 0, "dBD row stride must be a multiple of 8 and >= 2T-1");
  ESP_CHECK(ldctx % 8 == 0 && ldq % 8 == 0 && ldkv % 8 == 0 && ld_out % 8 == 0, "row strides must be multiples of 8");
  ESP_CHECK(((uintptr_t)dk & 15) == 0 && ((uintptr_t)dv & 15) == 0, "dk / dv must be 16-byte aligned");
  ESP_CHECK(B <= 65535 && H <= 65535, "grid limits");
  const long R = (long)B * T;
  esp_launch(attn_rowdot_kernel, (unsigned)((R + 7) / 8), 256, 0, st, (const bf16*)dctx, (const bf16*)ctx, R, T, H, (long)ldctx,
             rowdot_ws);
  ESP_LAUNCH_CHECK();
  CUtensorMap tdo, tqu, tv, tp, tpd;
  int rc;
  if ((rc = esp_make_tmap_bf16(&tdo, dctx, (long)H * kHd, T, ldctx, B, (long)T * ldctx, 1, 0, kTile))) return rc;
  if ((rc = esp_make_tmap_bf16(&tqu, qu, (long)H * kHd, T, ldq, B, (long)T * ldq, 1, 0, kTile))) return rc;
  if ((rc = esp_make_tmap_bf16(&tv, v, (long)H * kHd, T, ldkv, B, (long)T * ldkv, 1, 0, kTile))) return rc;
  if ((rc = esp_make_tmap_bf16(&tp, p, ldp_probs, T, ldp_probs, B, (long)T * ldp_probs, H, (long)B * T * ldp_probs, kTile))) return rc;
  if ((rc = esp_make_tmap_bf16(&tpd, pd, ldp_probs, T, ldp_probs, B, (long)T * ldp_probs, H, (long)B * T * ldp_probs, kTile))) return rc;
  Params pr;
  {
    const char* tl = getenv("ESP_ATTN_BWD_TIMELINE");
    pr.timeline = tl ? atoi(tl) : 0;  // bit 0: timeline; experiment switches: 2 no plain dS stores, 4 no skewed stores, 8 no skew loop
  }
  pr.B = B; pr.T = T; pr.H = H; pr.ld = ldp_probs; pr.ldp = ldbd;
  // the skewed copy comes from attn_skew_kernel below (ESP_ATTN_BWD_INLINE_SKEW=1: written by the fused kernel itself --
  // measured slower: its eight dS warps per SM cannot hide the latencies of the shifted stores)
  static int inline_skew = -1;
  if (inline_skew < 0) {
    const char* e = getenv("ESP_ATTN_BWD_INLINE_SKEW");
    inline_skew = (e && e[0] == '1') ? 1 : 0;
  }
  pr.D = rowdot_ws; pr.ds_out = (bf16*)ds; pr.dbd_out = inline_skew ? (bf16*)dbd : nullptr; pr.dk_out = (bf16*)dk; pr.dv_out = (bf16*)dv;
  pr.ld_out = ld_out;
  pr.drop_p = drop_p; pr.thresh = esp_dropout_thresh(drop_p); pr.seed = seed;
  pr.seed_ptr = (const unsigned long long*)seed_ptr;
  static bool configured = false;
  if (!configured) {
    ESP_CUDA(cudaFuncSetAttribute(attn_fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    configured = true;
  }
  dim3 grid((T + kTile - 1) / kTile, H, B);
  esp_launch(attn_fused_bwd_kernel, grid, kThreads, kSmemBytes, st, tdo, tqu, tv, tp, tpd, pr);
  ESP_LAUNCH_CHECK();
  int launches = 2;
  if (!inline_skew) {
    ESP_CHECK(T <= 8 * (kSkewMaxUnits - 2), "attention backward: at most %d frames", 8 * (kSkewMaxUnits - 2));
    const long rows = (long)H * B * T;
    esp_launch(attn_skew_kernel, (unsigned)((rows + kSkewWarps - 1) / kSkewWarps), kSkewWarps * 32, 0, st, (const bf16*)ds, (bf16*)dbd,
               rows, T, (int)ldp_probs, (int)ldbd);
    ESP_LAUNCH_CHECK();
    ++launches;
  }
  esp_count_launch(launches);
  return 0;
}
