// espresso_b200 -- single-query attention kernels for incremental (beam search) decoding.
//
// Replaces the incremental branch of fairseq/modules/multihead_attention.py:639-760,878-897 (self-attention over
// the cached prefix, encoder attention over per-sentence "beamable" K/V :661-669) and the O(t) index_select of
// every cached tensor per step (`reorder_incremental_state`, :964-989):
//   * K/V of step t are written once at cache[t, row, :] by the projection GEMM and never moved;
//   * beam reordering only re-gathers the int32 ancestry table anc[t][n] = cache row holding hypothesis n's
//     key/value at time t (esp_decode_update_ancestry), which the attention kernel follows.
// One warp per (hypothesis, head): phase 1 lanes <- keys (scores, fp32 softmax), phase 2 lanes <- head dims.
#include "common.cuh"
#include "espresso_b200.h"
#include <math.h>

void esp_count_launch(int n);

namespace {

constexpr int kMaxKeys = 1280;  // encoder frames after 4x subsampling (36 s -> 900) / decoder steps

// q [N, d]; K/V of hypothesis n at time t: kv + ((t*N + anc[t*N+n]) * 2d) (+d for V); T = step+1 keys.
__global__ void __launch_bounds__(128)
decode_self_attn_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kv, const int* __restrict__ anc, int N, int H,
                        int hd, int T, float scale, bf16* __restrict__ out) {
  extern __shared__ float sm[];  // [warps][T] probabilities
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
  if (gw >= N * H) return;
  const int n = gw / H, h = gw % H;
  const int d = H * hd;
  float* p = sm + warp * kMaxKeys;
  const bf16* qp = q + (long)n * d + h * hd;
  float mx = -INFINITY;
  for (int t = lane; t < T; t += 32) {
    const int row = anc[(long)t * N + n];
    const bf16* kp = kv + ((long)t * N + row) * 2 * d + h * hd;
    float s = 0.f;
    for (int c = 0; c < hd; c += 2) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(qp + c);
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(kp + c);
      s += __low2float(a) * __low2float(b) + __high2float(a) * __high2float(b);
    }
    s *= scale;
    p[t] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float e = __expf(p[t] - mx);
    p[t] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int c = lane * 2; c < hd; c += 64) {
    float a0 = 0.f, a1 = 0.f;
    for (int t = 0; t < T; ++t) {
      const int row = anc[(long)t * N + n];
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(kv + ((long)t * N + row) * 2 * d + d + h * hd + c);
      const float w = bf2f(f2bf(p[t] * inv));  // attention weights are cast to the model dtype before P.V
      a0 += w * __low2float(v);
      a1 += w * __high2float(v);
    }
    *reinterpret_cast<__nv_bfloat162*>(out + (long)n * d + h * hd + c) = __floats2bfloat162_rn(a0, a1);
  }
}

// q [N, d]; encoder K/V per SENTENCE: kv [bsz, Tk, 2d] (k | v); sentence of hypothesis n = n / beam.
__global__ void __launch_bounds__(128)
decode_cross_attn_kernel(const bf16* __restrict__ q, const bf16* __restrict__ kv, const int* __restrict__ lens, int N, int beam,
                         int H, int hd, int Tk, float scale, bf16* __restrict__ out) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
  if (gw >= N * H) return;
  const int n = gw / H, h = gw % H;
  const int d = H * hd;
  const int sent = n / beam;
  const int T = lens ? min(lens[sent], Tk) : Tk;  // encoder_padding_mask -> -inf
  float* p = sm + warp * kMaxKeys;
  const bf16* qp = q + (long)n * d + h * hd;
  const bf16* base = kv + (long)sent * Tk * 2 * d;
  float mx = -INFINITY;
  for (int t = lane; t < T; t += 32) {
    const bf16* kp = base + (long)t * 2 * d + h * hd;
    float s = 0.f;
    for (int c = 0; c < hd; c += 2) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(qp + c);
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(kp + c);
      s += __low2float(a) * __low2float(b) + __high2float(a) * __high2float(b);
    }
    s *= scale;
    p[t] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int t = lane; t < T; t += 32) {
    const float e = __expf(p[t] - mx);
    p[t] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int c = lane * 2; c < hd; c += 64) {
    float a0 = 0.f, a1 = 0.f;
    for (int t = 0; t < T; ++t) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(base + (long)t * 2 * d + d + h * hd + c);
      const float w = bf2f(f2bf(p[t] * inv));
      a0 += w * __low2float(v);
      a1 += w * __high2float(v);
    }
    *reinterpret_cast<__nv_bfloat162*>(out + (long)n * d + h * hd + c) = __floats2bfloat162_rn(a0, a1);
  }
}

// anc_out[t][n] = anc_in[t][new_order[n]] for t < step ; anc_out[step][n] = n
__global__ void __launch_bounds__(256)
update_ancestry_kernel(const int* __restrict__ anc_in, int* __restrict__ anc_out, const int* __restrict__ new_order, int N, int step) {
  const long total = (long)(step + 1) * N;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int t = (int)(i / N), n = (int)(i % N);
    anc_out[i] = (t == step) ? n : anc_in[(long)t * N + (new_order ? new_order[n] : n)];
  }
}

}  // namespace

extern "C" int esp_decode_self_attn(const void* q, const void* kv_cache, const int32_t* anc, int32_t N, int32_t H, int32_t hd,
                                    int32_t T, float scale, void* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ESP_CHECK(T >= 1 && T <= kMaxKeys && hd % 2 == 0, "decode self-attention: bad shape (T=%d)", T);
  if (N == 0) return 0;
  const int warps = 4;
  const size_t smem = (size_t)warps * kMaxKeys * sizeof(float);
  decode_self_attn_kernel<<<(N * H + warps - 1) / warps, warps * 32, smem, st>>>((const bf16*)q, (const bf16*)kv_cache, anc, N, H,
                                                                              hd, T, scale, (bf16*)out);
  ESP_LAUNCH_CHECK();
  esp_count_launch(1);
  return 0;
}

extern "C" int esp_decode_cross_attn(const void* q, const void* kv, const int32_t* lens, int32_t N, int32_t beam, int32_t H,
                                     int32_t hd, int32_t Tk, float scale, void* out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  ESP_CHECK(Tk >= 1 && Tk <= kMaxKeys && hd % 2 == 0 && beam >= 1, "decode cross-attention: bad shape (Tk=%d)", Tk);
  if (N == 0) return 0;
  const int warps = 4;
  const size_t smem = (size_t)warps * kMaxKeys * sizeof(float);
  decode_cross_attn_kernel<<<(N * H + warps - 1) / warps, warps * 32, smem, st>>>((const bf16*)q, (const bf16*)kv, lens, N, beam, H,
                                                                               hd, Tk, scale, (bf16*)out);
  ESP_LAUNCH_CHECK();
  esp_count_launch(1);
  return 0;
}

extern "C" int esp_decode_update_ancestry(const int32_t* anc_in, int32_t* anc_out, const int32_t* new_order, int32_t N,
                                          int32_t step, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) return 0;
  const long total = (long)(step + 1) * N;
  update_ancestry_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(anc_in, anc_out, new_order, N, step);
  ESP_LAUNCH_CHECK();
  esp_count_launch(1);
  return 0;
}
