// espresso_b200 -- CTC loss forward + backward fused with the fp32 log-softmax.
//
// Replaces (espresso/criterions/ctc_loss.py:59-103):
//   lprobs = model.get_normalized_probs(net_output, log_probs=True)   # fp32 [T',B,V] materialised
//   F.ctc_loss(lprobs, targets, in_lens, tgt_lens, blank, reduction="sum", zero_infinity)
// and its autograd backward through log_softmax.  The fp32 [T',B,V] log-prob tensor and its fp32
// gradient are never written: three kernels touch HBM with
//   prep : read logits once            -> row log-sum-exp + emissions gathered at the 2U+1 extended labels
//   scan : alpha (warp 0) and beta (warp 1) recursions per utterance, states in registers,
//          neighbour states by warp shuffle (no block barriers on the serial path)
//   grad : read logits again, write d(loss)/d(logits) = softmax - occupancy   (bf16)
// => algorithmic bytes 6*V per encoder frame.
#include "common.cuh"
#include "espresso_b200.h"
#include <math.h>

void esp_count_launch(int n);

namespace {

constexpr float kNegInf = -INFINITY;

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == kNegInf) return kNegInf;
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == kNegInf) return kNegInf;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : kNegInf;
  r = warp_max(r);
  __syncthreads();
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  __syncthreads();
  return r;
}

// ---- prep: row LSE + emissions at extended labels ------------------------------------------------
constexpr int kPrepThreads = 256;
constexpr int kPrepCache = 4;  // uint4 (8 bf16) per thread cached in registers: V <= 256*4*8 = 8192

__global__ void __launch_bounds__(kPrepThreads)
ctc_prep_kernel(const bf16* __restrict__ logits, long stride_b, long stride_t, int V, int t_max,
                const int* __restrict__ in_lens, const int* __restrict__ targets, int u_max,
                const int* __restrict__ tgt_lens, int blank, float* __restrict__ lse_out,
                float* __restrict__ lp_ext, int s_max) {
  __shared__ float red[32];
  const int b = blockIdx.y, t = blockIdx.x;
  if (t >= in_lens[b]) return;
  const bf16* row = logits + (long)b * stride_b + (long)t * stride_t;
  const bool vec_ok = ((((uintptr_t)row) & 15) == 0);
  const int nvec = vec_ok ? V / 8 : 0;
  uint4 cache[kPrepCache];
  float mx = kNegInf;
#pragma unroll
  for (int i = 0; i < kPrepCache; ++i) {
    const int vi = threadIdx.x + i * kPrepThreads;
    if (vi < nvec) {
      cache[i] = *reinterpret_cast<const uint4*>(row + vi * 8);
      float a, c;
      unpack_bf16x2(cache[i].x, a, c); mx = fmaxf(mx, fmaxf(a, c));
      unpack_bf16x2(cache[i].y, a, c); mx = fmaxf(mx, fmaxf(a, c));
      unpack_bf16x2(cache[i].z, a, c); mx = fmaxf(mx, fmaxf(a, c));
      unpack_bf16x2(cache[i].w, a, c); mx = fmaxf(mx, fmaxf(a, c));
    }
  }
  for (int vi = threadIdx.x + kPrepCache * kPrepThreads; vi < nvec; vi += kPrepThreads) {
    const uint4 q = *reinterpret_cast<const uint4*>(row + vi * 8);
    float a, c;
    unpack_bf16x2(q.x, a, c); mx = fmaxf(mx, fmaxf(a, c));
    unpack_bf16x2(q.y, a, c); mx = fmaxf(mx, fmaxf(a, c));
    unpack_bf16x2(q.z, a, c); mx = fmaxf(mx, fmaxf(a, c));
    unpack_bf16x2(q.w, a, c); mx = fmaxf(mx, fmaxf(a, c));
  }
  for (int v = nvec * 8 + threadIdx.x; v < V; v += kPrepThreads) mx = fmaxf(mx, bf2f(row[v]));
  mx = block_max(mx, red);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kPrepCache; ++i) {
    const int vi = threadIdx.x + i * kPrepThreads;
    if (vi < nvec) {
      float a, c;
      unpack_bf16x2(cache[i].x, a, c); sum += expf(a - mx) + expf(c - mx);
      unpack_bf16x2(cache[i].y, a, c); sum += expf(a - mx) + expf(c - mx);
      unpack_bf16x2(cache[i].z, a, c); sum += expf(a - mx) + expf(c - mx);
      unpack_bf16x2(cache[i].w, a, c); sum += expf(a - mx) + expf(c - mx);
    }
  }
  for (int vi = threadIdx.x + kPrepCache * kPrepThreads; vi < nvec; vi += kPrepThreads) {
    const uint4 q = *reinterpret_cast<const uint4*>(row + vi * 8);
    float a, c;
    unpack_bf16x2(q.x, a, c); sum += expf(a - mx) + expf(c - mx);
    unpack_bf16x2(q.y, a, c); sum += expf(a - mx) + expf(c - mx);
    unpack_bf16x2(q.z, a, c); sum += expf(a - mx) + expf(c - mx);
    unpack_bf16x2(q.w, a, c); sum += expf(a - mx) + expf(c - mx);
  }
  for (int v = nvec * 8 + threadIdx.x; v < V; v += kPrepThreads) sum += expf(bf2f(row[v]) - mx);
  sum = block_sum(sum, red);
  const float lse = mx + logf(sum);
  if (threadIdx.x == 0) lse_out[(long)b * t_max + t] = lse;
  const int U = tgt_lens[b];
  const int S = 2 * U + 1;
  float* dst = lp_ext + ((long)b * t_max + t) * s_max;
  for (int s = threadIdx.x; s < S; s += kPrepThreads) {
    const int lab = (s & 1) ? targets[(long)b * u_max + (s >> 1)] : blank;
    dst[s] = bf2f(row[lab]) - lse;
  }
}

// ---- scan: alpha / beta recursions, one warp each, C states per lane ----------------------------
template <int C>
__global__ void __launch_bounds__(64)
ctc_scan_kernel(const float* __restrict__ lp_ext, float* __restrict__ alpha, float* __restrict__ beta,
                int t_max, int s_max, const int* __restrict__ in_lens, const int* __restrict__ targets,
                int u_max, const int* __restrict__ tgt_lens, int blank, float* __restrict__ nll) {
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int T = in_lens[b];
  const int U = tgt_lens[b];
  const int S = 2 * U + 1;
  if (T <= 0) {
    if (threadIdx.x == 0) nll[b] = (U == 0) ? 0.f : INFINITY;
    return;
  }
  const long base = (long)b * t_max * s_max;
  const int s0 = lane * C;
  // which states may take the s-2 (alpha) / s+2 (beta) skip transition
  bool skip[C];
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const int s = s0 + i;
    skip[i] = false;
    if ((s & 1) && s < S) {
      const int u = s >> 1;
      if (warp == 0) {
        if (u >= 1) skip[i] = targets[(long)b * u_max + u] != targets[(long)b * u_max + u - 1];
      } else {
        if (u + 1 < U) skip[i] = targets[(long)b * u_max + u] != targets[(long)b * u_max + u + 1];
      }
    }
  }
  float cur[C], em[C], nxt[C];
  if (warp == 0) {
    // ---------------- alpha, t ascending ----------------
    const float* lp = lp_ext + base;
#pragma unroll
    for (int i = 0; i < C; ++i) {
      const int s = s0 + i;
      em[i] = (s < S) ? lp[s] : kNegInf;
      cur[i] = (s < 2 && s < S) ? em[i] : kNegInf;
    }
    float* al = alpha + base;
#pragma unroll
    for (int i = 0; i < C; ++i)
      if (s0 + i < S) al[s0 + i] = cur[i];
    if (T > 1) {
#pragma unroll
      for (int i = 0; i < C; ++i) nxt[i] = (s0 + i < S) ? lp[s_max + s0 + i] : kNegInf;
    }
    for (int t = 1; t < T; ++t) {
#pragma unroll
      for (int i = 0; i < C; ++i) em[i] = nxt[i];
      if (t + 1 < T) {
        const float* lpn = lp + (long)(t + 1) * s_max;
#pragma unroll
        for (int i = 0; i < C; ++i) nxt[i] = (s0 + i < S) ? lpn[s0 + i] : kNegInf;
      }
      // neighbours s-1, s-2 of this lane's first states live in the previous lane
      float p1 = __shfl_up_sync(0xffffffffu, cur[C - 1], 1);
      float p2 = __shfl_up_sync(0xffffffffu, cur[C - 2], 1);
      if (lane == 0) { p1 = kNegInf; p2 = kNegInf; }
      float nw[C];
#pragma unroll
      for (int i = 0; i < C; ++i) {
        // i==0: s-1 -> p1, s-2 -> p2 ; i==1: s-1 -> cur[0], s-2 -> p1
        const float b1 = (i == 0) ? p1 : cur[i >= 1 ? i - 1 : 0];
        const float b2 = (i == 0) ? p2 : (i == 1 ? p1 : cur[i >= 2 ? i - 2 : 0]);
        const float acc = skip[i] ? lse3(cur[i], b1, b2) : lse2(cur[i], b1);
        nw[i] = (s0 + i < S) ? acc + em[i] : kNegInf;
      }
      float* alt = al + (long)t * s_max;
#pragma unroll
      for (int i = 0; i < C; ++i) {
        cur[i] = nw[i];
        if (s0 + i < S) alt[s0 + i] = cur[i];
      }
    }
    // nll = -logsumexp(alpha_{T-1}(S-1), alpha_{T-1}(S-2))
    float last = kNegInf, last2 = kNegInf;
#pragma unroll
    for (int i = 0; i < C; ++i) {
      if (s0 + i == S - 1) last = cur[i];
      if (s0 + i == S - 2) last2 = cur[i];
    }
    last = warp_max(last);
    last2 = warp_max(last2);
    if (lane == 0) nll[b] = -lse2(last, last2);
  } else {
    // ---------------- beta, t descending ----------------
    const float* lp = lp_ext + base;
    const float* lpl = lp + (long)(T - 1) * s_max;
#pragma unroll
    for (int i = 0; i < C; ++i) {
      const int s = s0 + i;
      em[i] = (s < S) ? lpl[s] : kNegInf;
      cur[i] = (s < S && s >= S - 2) ? em[i] : kNegInf;
    }
    float* be = beta + base;
    {
      float* bt = be + (long)(T - 1) * s_max;
#pragma unroll
      for (int i = 0; i < C; ++i)
        if (s0 + i < S) bt[s0 + i] = cur[i];
    }
    if (T > 1) {
      const float* lpn = lp + (long)(T - 2) * s_max;
#pragma unroll
      for (int i = 0; i < C; ++i) nxt[i] = (s0 + i < S) ? lpn[s0 + i] : kNegInf;
    }
    for (int t = T - 2; t >= 0; --t) {
#pragma unroll
      for (int i = 0; i < C; ++i) em[i] = nxt[i];
      if (t - 1 >= 0) {
        const float* lpn = lp + (long)(t - 1) * s_max;
#pragma unroll
        for (int i = 0; i < C; ++i) nxt[i] = (s0 + i < S) ? lpn[s0 + i] : kNegInf;
      }
      float n1 = __shfl_down_sync(0xffffffffu, cur[0], 1);
      float n2 = __shfl_down_sync(0xffffffffu, cur[1], 1);
      if (lane == 31) { n1 = kNegInf; n2 = kNegInf; }
      float nw[C];
#pragma unroll
      for (int i = 0; i < C; ++i) {
        const float b1 = (i == C - 1) ? n1 : cur[i + 1 < C ? i + 1 : 0];
        const float b2 = (i == C - 1) ? n2 : (i == C - 2 ? n1 : cur[i + 2 < C ? i + 2 : 0]);
        const float acc = skip[i] ? lse3(cur[i], b1, b2) : lse2(cur[i], b1);
        nw[i] = (s0 + i < S) ? acc + em[i] : kNegInf;
      }
      float* bt = be + (long)t * s_max;
#pragma unroll
      for (int i = 0; i < C; ++i) {
        cur[i] = nw[i];
        if (s0 + i < S) bt[s0 + i] = cur[i];
      }
    }
  }
}

// ---- grad: softmax - occupancy ------------------------------------------------------------------
constexpr int kGradThreads = 256;

__global__ void __launch_bounds__(kGradThreads)
ctc_grad_kernel(const bf16* __restrict__ logits, long stride_b, long stride_t, int V, int ld_pad, int t_max,
                const int* __restrict__ in_lens, const int* __restrict__ targets, int u_max,
                const int* __restrict__ tgt_lens, int blank, int zero_infinity, float grad_scale,
                const float* __restrict__ lse_in, const float* __restrict__ lp_ext,
                const float* __restrict__ alpha, const float* __restrict__ beta, const float* __restrict__ nll,
                int s_max, bf16* __restrict__ grad) {
  extern __shared__ float occ[];
  const int b = blockIdx.y, t = blockIdx.x;
  bf16* grow = grad + (long)b * stride_b + (long)t * stride_t;
  const float loss = nll[b];
  const bool dead = (t >= in_lens[b]) || isinf(loss) || isnan(loss);
  if (dead) {
    for (int v = threadIdx.x; v < ld_pad; v += kGradThreads) grow[v] = f2bf(0.f);
    return;
  }
  for (int v = threadIdx.x; v < V; v += kGradThreads) occ[v] = 0.f;
  __syncthreads();
  const int S = 2 * tgt_lens[b] + 1;
  const long off = ((long)b * t_max + t) * s_max;
  for (int s = threadIdx.x; s < S; s += kGradThreads) {
    const float lab_lp = lp_ext[off + s];
    const float ab = alpha[off + s] + beta[off + s];
    if (ab != kNegInf) {
      const int lab = (s & 1) ? targets[(long)b * u_max + (s >> 1)] : blank;
      atomicAdd(&occ[lab], expf(ab - lab_lp + loss));
    }
  }
  __syncthreads();
  const bf16* row = logits + (long)b * stride_b + (long)t * stride_t;
  const float lse = lse_in[(long)b * t_max + t];
  for (int v = threadIdx.x; v < ld_pad; v += kGradThreads) {
    float g = 0.f;
    if (v < V) g = (expf(bf2f(row[v]) - lse) - occ[v]) * grad_scale;
    grow[v] = f2bf(g);
  }
}

__global__ void ctc_finalize_kernel(const float* __restrict__ nll, int B, int zero_infinity, float* __restrict__ loss) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float l = nll[b];
  if (zero_infinity && (isinf(l) || isnan(l))) l = 0.f;
  loss[b] = l;
}

inline long align256(long x) { return (x + 255) / 256 * 256; }

}  // namespace

extern "C" int64_t esp_ctc_workspace_bytes(int32_t B, int32_t t_max, int32_t u_max) {
  const long s_max = 2L * u_max + 1;
  const long cells = (long)B * t_max * s_max;
  return align256((long)B * t_max * 4) + 3 * align256(cells * 4) + align256((long)B * 4);
}

extern "C" int esp_ctc_loss(const void* logits, int64_t stride_b, int64_t stride_t, int32_t V, int32_t B,
                            int32_t t_max, const int32_t* in_lens, const int32_t* targets, int32_t u_max,
                            const int32_t* tgt_lens, int32_t blank, int32_t zero_infinity, float grad_scale,
                            float* loss, void* grad, void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ESP_CHECK(B >= 0 && t_max >= 0 && u_max >= 0 && V > 0, "bad CTC shape");
  if (B == 0) return 0;
  ESP_CHECK(logits && in_lens && tgt_lens && loss && workspace, "null pointer passed to esp_ctc_loss");
  ESP_CHECK(u_max == 0 || targets != nullptr, "targets is null");
  ESP_CHECK(blank >= 0 && blank < V, "blank index out of range");
  const int s_max = 2 * u_max + 1;
  ESP_CHECK(s_max <= 32 * 32, "CTC target too long for the register scan (u_max=%d > 511)", u_max);
  char* ws = (char*)workspace;
  float* lse = (float*)ws; ws += align256((long)B * t_max * 4);
  const long cells = (long)B * t_max * s_max;
  float* lp_ext = (float*)ws; ws += align256(cells * 4);
  float* alpha = (float*)ws; ws += align256(cells * 4);
  float* beta = (float*)ws; ws += align256(cells * 4);
  float* nll = (float*)ws;
  if (t_max > 0) {
    ctc_prep_kernel<<<dim3(t_max, B), kPrepThreads, 0, st>>>((const bf16*)logits, stride_b, stride_t, V, t_max, in_lens,
                                                            targets, u_max, tgt_lens, blank, lse, lp_ext, s_max);
    ESP_LAUNCH_CHECK();
  }
#define ESP_SCAN(C)                                                                                          \
  ctc_scan_kernel<C><<<B, 64, 0, st>>>(lp_ext, alpha, beta, t_max, s_max, in_lens, targets, u_max, tgt_lens, \
                                       blank, nll)
  if (s_max <= 64) ESP_SCAN(2);
  else if (s_max <= 128) ESP_SCAN(4);
  else if (s_max <= 256) ESP_SCAN(8);
  else if (s_max <= 512) ESP_SCAN(16);
  else ESP_SCAN(32);
#undef ESP_SCAN
  ESP_LAUNCH_CHECK();
  ctc_finalize_kernel<<<(B + 127) / 128, 128, 0, st>>>(nll, B, zero_infinity, loss);
  ESP_LAUNCH_CHECK();
  int launches = 3;
  if (grad && t_max > 0) {
    // padded row width: rows are written up to the next multiple of 8 when the row stride leaves room
    long min_stride = stride_t < stride_b ? stride_t : stride_b;
    int ld_pad = V;
    if (B == 1 && t_max > 1) min_stride = stride_t;
    if (t_max == 1 && B > 1) min_stride = stride_b;
    const int v8 = (V + 7) / 8 * 8;
    if (min_stride >= v8) ld_pad = v8;
    ESP_CHECK((size_t)V * 4 <= 200 * 1024, "vocabulary too large for the CTC gradient kernel");
    static int cfg_bytes = 0;
    if ((int)(V * 4) > cfg_bytes) {
      ESP_CUDA(cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V * 4));
      cfg_bytes = V * 4;
    }
    ctc_grad_kernel<<<dim3(t_max, B), kGradThreads, V * 4, st>>>(
        (const bf16*)logits, stride_b, stride_t, V, ld_pad, t_max, in_lens, targets, u_max, tgt_lens, blank,
        zero_infinity, grad_scale, lse, lp_ext, alpha, beta, nll, s_max, (bf16*)grad);
    ESP_LAUNCH_CHECK();
    ++launches;
  }
  esp_count_launch(launches);
  return 0;
}
