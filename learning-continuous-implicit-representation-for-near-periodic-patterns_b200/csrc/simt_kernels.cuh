// CUDA-core kernels of the NPP-Net train step that are HBM/latency bound rather than
// tensor bound: the periodicity-aware positional encoding, the 256->3 RGB head, the
// masked-MSE loss, gradient finalisation, Adam and the fp16 shadow-weight refresh.
// Every kernel cites the reference lines whose arithmetic it reproduces.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <math_constants.h>
#include "grad_scale.cuh"
#include "ptx_sm100.cuh"

namespace npp {

constexpr int MAX_TOPK = 8;   // proposals
constexpr int MAX_AUG = 16;   // (freq_scale x freq_offset x angle_offset) combinations per direction
constexpr int MAX_FREQ = 16;  // Fourier frequencies of the NeRF-style expansion

// Encoder constants, computed on the host in fp32 with the reference's own torch ops
// (models/embedder.py:112-127) so that cos(theta), sin(theta) and the period match bit for bit.
struct EncTable {
  int topk;
  int n_aug;
  int n_freq;
  int include_input;  // 1: [x_norm, sin/cos...] per direction (embedder.py:105-109)
  float res_h, res_w;
  float cos_t[MAX_TOPK][2][MAX_AUG];
  float sin_t[MAX_TOPK][2][MAX_AUG];
  float period[MAX_TOPK][2][MAX_AUG];
  float freq[MAX_FREQ];
};

// torch.remainder semantics (sign of the divisor), embedder.py:127 uses the % operator.
// fmodf is exact (no rounding): for a moderate quotient it is one division, one truncation and one FMA.
// q' = trunc(fl(a / b)) is the true quotient or one past it (rounding is monotonic), a - q' b is exactly
// representable in both cases, so the FMA returns it exactly and a single correction step restores the sign rule.
__device__ __forceinline__ float npp_fmodf(float a, float b) {
  const float q = truncf(__fdiv_rn(a, b));
  if (!(b > 0.0f) || !(fabsf(q) < 1048576.0f)) return fmodf(a, b);
  float r = fmaf(-q, b, a);
  if (a >= 0.0f) {
    if (r < 0.0f) r += b;
  } else {
    if (r > 0.0f) r -= b;
  }
  return r == 0.0f ? copysignf(0.0f, a) : r;
}

__device__ __forceinline__ float torch_remainder(float a, float b) {
  float r = npp_fmodf(a, b);
  if (r != 0.0f && ((r < 0.0f) != (b < 0.0f))) r += b;
  return r;
}

// Phase phi of augmentation `aug` of direction `dir` (embedder.py:127-130):
// (((y*cos + x*sin) % P) / P) * 2 * pi, every step rounded to fp32 like the eager torch ops.
__device__ __forceinline__ float npp_phase(const EncTable& t, int j, int dir, int aug, float y, float x) {
  const float ct = t.cos_t[j][dir][aug], st = t.sin_t[j][dir][aug], P = t.period[j][dir][aug];
  const float proj = __fadd_rn(__fmul_rn(y, ct), __fmul_rn(x, st));
  const float frac = __fdiv_rn(torch_remainder(proj, P), P);
  return __fmul_rn(__fmul_rn(frac, 2.0f), 3.14159265358979323846f);
}

// (x / res[1] - 0.5) * 2  |  (y / res[0] - 0.5) * 2      embedder.py:107-108
__device__ __forceinline__ float npp_norm_coord(const EncTable& t, int dir, float y, float x) {
  const float v = dir == 0 ? __fdiv_rn(x, t.res_w) : __fdiv_rn(y, t.res_h);
  return __fmul_rn(__fsub_rn(v, 0.5f), 2.0f);
}

// Base periodic feature c (0 .. 2*(1+2*n_aug)-1) of one proposal for pixel (row y, col x).
// Layout (embedder.py:140-148): fn_x list then fn_y list; each list is
// [normalised coordinate, sin phi_0, cos phi_0, sin phi_1, cos phi_1, ...].
__device__ __forceinline__ float npp_base_feature(const EncTable& t, int j, int c, float y, float x) {
  const int per_dir = t.include_input + 2 * t.n_aug;
  const int dir = c / per_dir;
  int r = c - dir * per_dir;
  if (t.include_input) {
    if (r == 0) return npp_norm_coord(t, dir, y, x);
    r -= 1;
  }
  const float phi = npp_phase(t, j, dir, r >> 1, y, x);
  return (r & 1) ? cosf(phi) : sinf(phi);
}

// Expanded encoding of `rows` rows of one proposal -> fp16, reference column order
// out[:, b*B + c]: b = 0 identity, b = 1+2k sin(f_k u_c), b = 2+2k cos(f_k u_c)   (embedder.py:41-44,56)
// Three phases per block: (0) the B base features of every row in fp32 (one phase evaluation per sin/cos pair),
// (1) the Fourier expansion, one thread per (row, pair of adjacent base features) so that every shared-memory store
// is a half2, (2) copy-out with 16-byte stores: the tile row is kept at the same offset modulo 16 bytes as its
// destination row, so whole aligned chunks move as uint4 and only the two ragged ends fall back to half2.
constexpr int ENC_THREADS = 256;
__host__ __device__ inline int enc_row_stride(int width) { return (width + 8 + 7) & ~7; }  // halfs, room for the shift
__host__ __device__ inline int enc_base_bytes(int rows, int B) { return (rows * B * 4 + 15) & ~15; }
__global__ void __launch_bounds__(ENC_THREADS) npp_encode_kernel(const float* __restrict__ coords, int n, EncTable t,
                                                                 __half* __restrict__ enc1, int ld1,
                                                                 __half* __restrict__ enca, int lda, int rows,
                                                                 float* __restrict__ zero_a, int zero_a_n,
                                                                 float* __restrict__ zero_b) {
  extern __shared__ __align__(16) uint8_t enc_smem[];
  // first kernel of a fused train step: also clears the step's accumulators (bias-gradient sums, head gradients,
  // max|g|) and the loss scalar, which saves two memset nodes per step
  if (blockIdx.x == 0 && blockIdx.y == 0) {
    for (int i = threadIdx.x; i < zero_a_n; i += blockDim.x) zero_a[i] = 0.f;
    if (zero_b != nullptr && threadIdx.x == 0) *zero_b = 0.f;
  }
  const int per_dir = t.include_input + 2 * t.n_aug;
  const int B = 2 * per_dir;
  const int F = 1 + 2 * t.n_freq;
  const int width = B * F;
  const int RS = enc_row_stride(width);
  float* ubase = reinterpret_cast<float*>(enc_smem);                              // [rows][B] fp32 base features
  __half* tile = reinterpret_cast<__half*>(enc_smem + enc_base_bytes(rows, B));   // [rows][RS] fp16
  const int j = blockIdx.y;
  const int row0 = blockIdx.x * rows;
  __half* dst = j == 0 ? enc1 : enca + (size_t)(j - 1) * width;
  const int ld = j == 0 ? ld1 : lda;

  // index decomposition by a small runtime divisor d: (idx + 0.5) / d is at least 0.5 / d away from an integer,
  // far more than the fp32 error of the product for idx < 2^16, d <= 64
  const int per_dir_items = t.include_input + t.n_aug;
  const int items_row = 2 * per_dir_items;
  const float inv_items_row = 1.0f / (float)items_row, inv_per_dir_items = 1.0f / (float)per_dir_items;
  for (int idx = threadIdx.x; idx < rows * items_row; idx += blockDim.x) {
    const int r = __float2int_rz(((float)idx + 0.5f) * inv_items_row), q = idx - r * items_row;
    const int dir = __float2int_rz(((float)q + 0.5f) * inv_per_dir_items), a = q - dir * per_dir_items;
    const int row = row0 + r;
    if (row >= n) continue;
    const float y = coords[2 * row], x = coords[2 * row + 1];
    float* u = ubase + r * B + dir * per_dir;
    if (t.include_input && a == 0) {
      u[0] = npp_norm_coord(t, dir, y, x);
    } else {
      const int aug = a - t.include_input;
      float sn, cs;
      sincosf(npp_phase(t, j, dir, aug, y, x), &sn, &cs);
      u[t.include_input + 2 * aug] = sn;
      u[t.include_input + 2 * aug + 1] = cs;
    }
  }
  __syncthreads();

  const int half_b = B >> 1;
  const float inv_half_b = 1.0f / (float)half_b;
  const int shift0 = static_cast<int>((reinterpret_cast<uintptr_t>(dst) >> 1) & 7);  // every row's, if ld % 8 == 0
  for (int idx = threadIdx.x; idx < rows * half_b; idx += blockDim.x) {
    const int r = __float2int_rz(((float)idx + 0.5f) * inv_half_b), cp = idx - r * half_b;
    const int row = row0 + r;
    if (row >= n) continue;
    const int shift = (ld & 7) == 0 ? shift0 : static_cast<int>((reinterpret_cast<uintptr_t>(dst + (size_t)row * ld) >> 1) & 7);
    const float2 u = *reinterpret_cast<const float2*>(ubase + r * B + 2 * cp);
    __half2* o = reinterpret_cast<__half2*>(tile + r * RS + shift + 2 * cp);
    o[0] = __floats2half2_rn(u.x, u.y);
    o += half_b;
    for (int k = 0; k < t.n_freq; ++k) {
      const float f = t.freq[k];
      const float a0 = __fmul_rn(u.x, f), a1 = __fmul_rn(u.y, f);  // p_fn(x * freq), embedder.py:43
      o[0] = __floats2half2_rn(__sinf(a0), __sinf(a1));
      o[half_b] = __floats2half2_rn(__cosf(a0), __cosf(a1));
      o += B;
    }
  }
  __syncthreads();

  if ((ld & 7) == 0) {
    // Row pitch is a multiple of 16 bytes (every plan buffer): all rows of the block share one alignment shift, so
    // thread t owns 16-byte chunk column t % 64 and walks rows t / 64, +4, ... with plain pointer increments.
    const int shift = static_cast<int>((reinterpret_cast<uintptr_t>(dst) >> 1) & 7);
    const int end = shift + width;
    const int first_full = (shift + 7) >> 3, last_full = end >> 3;
    const int nch = (end + 7) >> 3;
    const int rows_here = min(rows, n - row0);
    for (int ch = threadIdx.x & 63; ch < nch; ch += 64) {
      const bool whole = ch >= first_full && ch < last_full;
      const uint8_t* sp = reinterpret_cast<const uint8_t*>(tile) + (size_t)(threadIdx.x >> 6) * RS * 2 + ch * 16;
      uint8_t* gp = reinterpret_cast<uint8_t*>(dst + (size_t)(row0 + (threadIdx.x >> 6)) * ld - shift) + ch * 16;
      const size_t sstep = (size_t)RS * 2 * (ENC_THREADS / 64), gstep = (size_t)ld * 2 * (ENC_THREADS / 64);
      for (int r = threadIdx.x >> 6; r < rows_here; r += ENC_THREADS / 64, sp += sstep, gp += gstep) {
        if (whole) {
          *reinterpret_cast<uint4*>(gp) = *reinterpret_cast<const uint4*>(sp);
        } else {  // ragged first / last chunk: the half2 pieces that belong to this proposal
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int pos = ch * 8 + 2 * i;
            if (pos >= shift && pos < end)
              *reinterpret_cast<__half2*>(gp + 4 * i) = *reinterpret_cast<const __half2*>(sp + 4 * i);
          }
        }
      }
    }
    return;
  }
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = wib; r < rows; r += nw) {
    const int row = row0 + r;
    if (row >= n) break;
    __half* grow = dst + (size_t)row * ld;
    const int shift = static_cast<int>((reinterpret_cast<uintptr_t>(grow) >> 1) & 7);
    const __half* srow = tile + r * RS;
    __half* gal = grow - shift;  // 16-byte aligned
    const int end = shift + width;
    const int first_full = (shift + 7) >> 3, last_full = end >> 3;  // 8-half chunks [first_full, last_full) are whole
    for (int ch = first_full + lane; ch < last_full; ch += 32)
      reinterpret_cast<uint4*>(gal)[ch] = reinterpret_cast<const uint4*>(srow)[ch];
    if (lane < 8) {  // ragged ends: at most three half2 in front and three behind
      const int head_end = min(first_full << 3, end);
      const int pos = lane < 4 ? shift + 2 * lane : (last_full << 3) + 2 * (lane - 4);
      const bool ok = lane < 4 ? pos < head_end : (last_full >= first_full && pos < end);
      if (ok) *reinterpret_cast<__half2*>(gal + pos) = *reinterpret_cast<const __half2*>(srow + pos);
    }
  }
}

// fp32 materialised encoding in the reference layout [n, topk*width] (parity tests and
// the "materialised mode" of the drop-in embedder).
__global__ void npp_encode_f32_kernel(const float* __restrict__ coords, int n, EncTable t, float* __restrict__ out) {
  const int B = 2 * (t.include_input + 2 * t.n_aug);
  const int F = 1 + 2 * t.n_freq;
  const int width = B * F;
  const long long total = (long long)n * t.topk * B;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = idx % B;
    const int j = (idx / B) % t.topk;
    const long long row = idx / ((long long)B * t.topk);
    const float y = coords[2 * row], x = coords[2 * row + 1];
    const float u = npp_base_feature(t, j, c, y, x);
    float* o = out + row * (long long)(t.topk * width) + j * width + c;
    o[0] = u;
    for (int k = 0; k < t.n_freq; ++k) {
      const float a = __fmul_rn(u, t.freq[k]);
      o[(1 + 2 * k) * B] = sinf(a);
      o[(2 + 2 * k) * B] = cosf(a);
    }
  }
}

// ---------------------------------------------------------------- search mode (NPP_Net_light)
// create_npp_net(is_search=True) (models/helpers.py:87-103) builds two encoders:
//   Embedder_periodic without raw input and without the Fourier expansion (embedder.py:93-95): 4*n_aug columns,
//     fn_x list then fn_y list, each [sin phi_0, cos phi_0, sin phi_1, ...]  (npp_base_feature with include_input 0);
//   the 2-D Embedder on coordinates normalised in place, row first (embedder.py:51-56,76-80): 2 + 4*n_freq columns
//     [yn, xn, sin(f_0 yn), sin(f_0 xn), cos(f_0 yn), cos(f_0 xn), sin(f_1 yn), ...].
// Work item = (row, group): groups [0, 2*n_aug) are the (sin, cos) pairs of one phase, group 2*n_aug is the raw
// normalised pair, the following n_freq groups are the four values of one frequency.
__device__ __forceinline__ float npp_search_norm(const EncTable& t, int d, float y, float x) {
  // ((inputs[:, d] / res[d]) - 0.5) * 2   embedder.py:53-54 (d = 0: row / H, d = 1: col / W)
  const float v = d == 0 ? __fdiv_rn(y, t.res_h) : __fdiv_rn(x, t.res_w);
  return __fmul_rn(__fsub_rn(v, 0.5f), 2.0f);
}

template <typename T>
__device__ __forceinline__ void npp_search_store2(T* p, float a, float b);
template <>
__device__ __forceinline__ void npp_search_store2<__half>(__half* p, float a, float b) {
  *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b);
}
template <>
__device__ __forceinline__ void npp_search_store2<float>(float* p, float a, float b) {
  p[0] = a;
  p[1] = b;
}

template <typename T>
__device__ __forceinline__ void npp_search_item(const EncTable& t, int g, float y, float x, T* per, T* pos) {
  const int n_pairs = 2 * t.n_aug;
  if (g < n_pairs) {
    const int dir = g / t.n_aug, aug = g - dir * t.n_aug;
    const float phi = npp_phase(t, 0, dir, aug, y, x);
    npp_search_store2<T>(per + 2 * g, sinf(phi), cosf(phi));
  } else {
    const float yn = npp_search_norm(t, 0, y, x), xn = npp_search_norm(t, 1, y, x);
    const int k = g - n_pairs - 1;
    if (k < 0) {
      npp_search_store2<T>(pos, yn, xn);
    } else {
      const float ay = __fmul_rn(yn, t.freq[k]), ax = __fmul_rn(xn, t.freq[k]);
      npp_search_store2<T>(pos + 2 + 4 * k, sinf(ay), sinf(ax));
      npp_search_store2<T>(pos + 4 + 4 * k, cosf(ay), cosf(ax));
    }
  }
}

// fp16 operand buffers of the fused path: per = enc1 [n, ld1], pos = the second K segment of pos_linears.0 [n, ldp].
__global__ void __launch_bounds__(256) npp_encode_search_kernel(const float* __restrict__ coords, int n, EncTable t,
                                                                __half* __restrict__ enc1, int ld1,
                                                                __half* __restrict__ pos, int ldp,
                                                                float* __restrict__ zero_a, int zero_a_n,
                                                                float* __restrict__ zero_b,
                                                                const int* __restrict__ step, int step_off) {
  // step != nullptr (npp_fit_run's re-launched step graph): batch *step (+ step_off: encoded one iteration ahead) of a
  // [iters, n, 2] array, loss slot *step
  if (step != nullptr) {
    const int sidx = *step + step_off;
    coords += (size_t)sidx * n * 2;
    if (zero_b != nullptr) zero_b += sidx;
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < zero_a_n; i += blockDim.x) zero_a[i] = 0.f;
    if (zero_b != nullptr && threadIdx.x == 0) *zero_b = 0.f;
  }
  const int groups = 2 * t.n_aug + 1 + t.n_freq;
  const long long total = (long long)n * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / groups;
    const int g = (int)(idx - row * groups);
    npp_search_item<__half>(t, g, coords[2 * row], coords[2 * row + 1], enc1 + row * ld1, pos + row * ldp);
  }
}

// fp32 materialised search encodings: out [n, 4*n_aug + 2 + 4*n_freq] = [periodic | positional].
__global__ void npp_encode_search_f32_kernel(const float* __restrict__ coords, int n, EncTable t,
                                             float* __restrict__ out) {
  const int groups = 2 * t.n_aug + 1 + t.n_freq;
  const int wper = 4 * t.n_aug, width = wper + 2 + 4 * t.n_freq;
  const long long total = (long long)n * groups;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / groups;
    const int g = (int)(idx - row * groups);
    float* o = out + row * width;
    npp_search_item<float>(t, g, coords[2 * row], coords[2 * row + 1], o, o + wper);
  }
}

// Materialised mode: a caller-supplied fp32 encoding [n, topk*width] (reference layout, e.g. rows gathered
// from the reference's precomputed table, NPP_completion/train.py:166-181) -> the fp16 operand buffers.
__global__ void npp_load_encoding_kernel(const float* __restrict__ enc, int n, int cols, int width,
                                         __half* __restrict__ enc1, int ld1, __half* __restrict__ enca, int lda) {
  const long long total = (long long)n * cols;   // columns [0, width) feed enc1, [width, cols) the second buffer
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / cols;
    const int c = (int)(idx - row * cols);
    const __half v = __float2half_rn(enc[idx]);
    if (c < width) enc1[row * ld1 + c] = v;
    else enca[row * lda + (c - width)] = v;
  }
}

// ------------------------------------------------------------------ RGB head
// logits = h_P . W_rgb^T + b_rgb   (models/networks.py:94), one warp per row.
__global__ void __launch_bounds__(256) npp_head_fwd_kernel(const __half* __restrict__ hp, int ld, int width, int n,
                                                           const float* __restrict__ w, const float* __restrict__ b,
                                                           float* __restrict__ logits) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const __half* h = hp + (size_t)warp * ld;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int k = lane * 8; k < width; k += 256) {
    const uint4 raw = *reinterpret_cast<const uint4*>(h + k);
    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h2[i]);
      const int kk = k + 2 * i;
      a0 = fmaf(f.x, w[kk], fmaf(f.y, w[kk + 1], a0));
      a1 = fmaf(f.x, w[width + kk], fmaf(f.y, w[width + kk + 1], a1));
      a2 = fmaf(f.x, w[2 * width + kk], fmaf(f.y, w[2 * width + kk + 1], a2));
    }
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, s);
    a1 += __shfl_xor_sync(0xffffffffu, a1, s);
    a2 += __shfl_xor_sync(0xffffffffu, a2, s);
  }
  if (lane == 0) {
    logits[3 * (size_t)warp + 0] = a0 + b[0];
    logits[3 * (size_t)warp + 1] = a1 + b[1];
    logits[3 * (size_t)warp + 2] = a2 + b[2];
  }
}

// Inference: RGB head + squashing (models/helpers.py:55-58: sigmoid for normalize_type 1, tanh for 2), written straight
// into an [H, W, 3] image at the pixel each coordinate row names -- the scatter of NPP_completion/train.py:288-296
// (pred_img[:, coord[:, 0], coord[:, 1], :] = pred) without the intermediate [n, 3] tensor.  One warp per row.
__global__ void __launch_bounds__(256) npp_head_render_kernel(const __half* __restrict__ hp, int ld, int width, int n,
                                                              const float* __restrict__ w, const float* __restrict__ b,
                                                              const float* __restrict__ coords,
                                                              float* __restrict__ image, int img_h, int img_w,
                                                              int normalize_type) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const __half* h = hp + (size_t)warp * ld;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int k = lane * 8; k < width; k += 256) {
    const uint4 raw = *reinterpret_cast<const uint4*>(h + k);
    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h2[i]);
      const int kk = k + 2 * i;
      a0 = fmaf(f.x, w[kk], fmaf(f.y, w[kk + 1], a0));
      a1 = fmaf(f.x, w[width + kk], fmaf(f.y, w[width + kk + 1], a1));
      a2 = fmaf(f.x, w[2 * width + kk], fmaf(f.y, w[2 * width + kk + 1], a2));
    }
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, s);
    a1 += __shfl_xor_sync(0xffffffffu, a1, s);
    a2 += __shfl_xor_sync(0xffffffffu, a2, s);
  }
  if (lane < 3) {
    const int y = (int)coords[2 * (size_t)warp], x = (int)coords[2 * (size_t)warp + 1];
    if (y < 0 || y >= img_h || x < 0 || x >= img_w) return;
    const float z = (lane == 0 ? a0 : (lane == 1 ? a1 : a2)) + b[lane];
    const float v = normalize_type == 2 ? tanhf(z) : 1.0f / (1.0f + expf(-z));
    image[((size_t)y * img_w + x) * 3 + lane] = v;
  }
}

// Patch crops of the sampler (models/sampler.py:262-296 via utils/extract_glimpse.py with mode='nearest',
// padding_mode='zeros'): out[m, c, i, j] = img[rows[m, i], cols[m, j], c], zero where an index falls outside the image.
// img is the [H, W, C] image as the train scripts hold it; rows / cols are the per-window index tables (contiguous for
// integer centroids, grid_sample's nearest rounding for fractional ones).  Pure data movement: bit-exact.
__global__ void __launch_bounds__(256) npp_gather_windows_kernel(const float* __restrict__ img, int H, int W, int C,
                                                                 const long long* __restrict__ rows,
                                                                 const long long* __restrict__ cols, int M, int h, int w,
                                                                 float* __restrict__ out) {
  const long long total = (long long)M * C * h * w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % w);
    const int i = (int)((idx / w) % h);
    const int c = (int)((idx / ((long long)w * h)) % C);
    const int m = (int)(idx / ((long long)w * h * C));
    const long long r = rows[(long long)m * h + i], q = cols[(long long)m * w + j];
    float v = 0.f;
    if (r >= 0 && r < H && q >= 0 && q < W) v = img[((size_t)r * W + (size_t)q) * C + c];
    out[idx] = v;
  }
}

// Candidate real-patch centroids of GridPatchSampler.sample_patch_real (models/sampler.py:127-216, integer lattice):
// for fake centroid n and (i, j) in [-10, 10)^2 (q = (i + 10) * 20 + (j + 10), the order of the reference's
// meshgrid / reshape), c = cent[n] + i * s1 + j * s2 in (row, col).  keep[n * 400 + q] = 1 iff the centroid lies
// strictly inside the image (sampler.py:156-159) and its window [row - hh, row + hh) x [col - wh, col + wh) holds at
// most `thresh` unknown pixels (mask < 0.5; outside the image counts as unknown: zero padding), counted in O(1) from
// the summed-area table sat [H + 1, W + 1] (the reference crops every candidate window, sampler.py:171-190).
__global__ void __launch_bounds__(256) npp_sampler_candidates_kernel(const long long* __restrict__ sat, int H, int W,
                                                                     const long long* __restrict__ cent, int n_samples,
                                                                     long long s1r, long long s1c, long long s2r,
                                                                     long long s2c, int hh, int wh, float thresh,
                                                                     unsigned char* __restrict__ keep) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_samples * 400) return;
  const int n = idx / 400, q = idx - n * 400;
  const long long i = q / 20 - 10, j = q % 20 - 10;
  const long long r = cent[2 * n] + i * s1r + j * s2r, c = cent[2 * n + 1] + i * s1c + j * s2c;
  const bool in_bound = r > 0 && r < H - 1 && c > 0 && c < W - 1;
  const long long h = 2 * hh, w = 2 * wh, r0 = r - hh, c0 = c - wh;
  const long long ra = min(max(r0, 0LL), (long long)H), rb = min(max(r0 + h, 0LL), (long long)H);
  const long long ca = min(max(c0, 0LL), (long long)W), cb = min(max(c0 + w, 0LL), (long long)W);
  const long long pitch = W + 1;
  const long long inside = sat[rb * pitch + cb] - sat[ra * pitch + cb] - sat[rb * pitch + ca] + sat[ra * pitch + ca];
  const long long unknown = inside + (h * w - (rb - ra) * (cb - ca));
  keep[idx] = (in_bound && !((float)unknown > thresh)) ? 1 : 0;
}

// sigmoid + masked MSE (models/helpers.py:55-56, models/mse_calculator.py:13-27 'l2' branch)
//   y_hat = sigmoid(logit); d = (y_hat - y) * (m + 0.3 (1 - m)); loss = mean(d^2) over n_norm*3
// also emits g = dLoss/dlogit and the running max |g| (as uint bits) for the fp16 delta scale.
__global__ void __launch_bounds__(256) npp_mse_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                      const float* __restrict__ mask, int n, float inv_count,
                                                      float* __restrict__ pred, float* __restrict__ g,
                                                      float* __restrict__ loss, unsigned int* __restrict__ amax_bits) {
  float lsum = 0.f, lmax = 0.f;
  const int total = n * 3;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int row = idx / 3;
    const float z = logits[idx];
    const float yh = 1.0f / (1.0f + expf(-z));
    const float m = mask ? mask[row] : 1.0f;
    const float w = m + (1.0f - m) * 0.3f;
    const float d = (yh - target[idx]) * w;
    lsum += d * d;
    const float gi = 2.0f * d * w * inv_count * yh * (1.0f - yh);
    if (pred) pred[idx] = yh;
    g[idx] = gi;
    lmax = fmaxf(lmax, fabsf(gi));
  }
  __shared__ float ssum[8], smax[8];
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    lsum += __shfl_xor_sync(0xffffffffu, lsum, s);
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, s));
  }
  if ((threadIdx.x & 31) == 0) {
    ssum[threadIdx.x >> 5] = lsum;
    smax[threadIdx.x >> 5] = lmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
      a += ssum[i];
      b = fmaxf(b, smax[i]);
    }
    atomicAdd(loss, a * inv_count);
    atomicMax(amax_bits, __float_as_uint(b));  // non-negative floats order like their bit patterns
  }
}

// Fused RGB head + sigmoid + masked MSE for the train-step path: one warp per row computes the three
// logits (networks.py:94), lanes 0..2 then apply helpers.py:55-56 and mse_calculator.py:13-27 ('l2').
__device__ __forceinline__ void npp_head_loss_part(const __half* __restrict__ hp, int ld, int width, int n,
                                                   const float* __restrict__ w, const float* __restrict__ b,
                                                   const float* __restrict__ target, const float* __restrict__ mask,
                                                   float inv_count, float* logits, float* g, float* loss,
                                                   unsigned int* amax_bits) {
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  float lsum = 0.f, lmax = 0.f;
  // width == 256: lane owns columns [8*lane, 8*lane+8); their 3x8 head weights stay in registers
  float wr[3][8];
  const bool fast = width == 256;
  if (fast) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) wr[c][i] = w[c * width + lane * 8 + i];
  }
  const int wstride = gridDim.x * 8;
  for (int row_base = blockIdx.x * 8 + wib; row_base < n; row_base += 4 * wstride) {
    // up to 4 rows per iteration so that four 512-byte row loads are in flight per warp
    float a[4][3];
    uint4 raw[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = row_base + u * wstride;
      raw[u] = make_uint4(0, 0, 0, 0);
      if (fast && row < n) raw[u] = *reinterpret_cast<const uint4*>(hp + (size_t)row * ld + lane * 8);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int row = row_base + u * wstride;
      a[u][0] = a[u][1] = a[u][2] = 0.f;
      if (row >= n) continue;
      if (fast) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[u]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __half22float2(h2[i]);
#pragma unroll
          for (int c = 0; c < 3; ++c) a[u][c] = fmaf(f.x, wr[c][2 * i], fmaf(f.y, wr[c][2 * i + 1], a[u][c]));
        }
      } else {
        const __half* h = hp + (size_t)row * ld;
        for (int k = lane * 8; k < width; k += 256) {
          const uint4 rw = *reinterpret_cast<const uint4*>(h + k);
          const __half2* h2 = reinterpret_cast<const __half2*>(&rw);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(h2[i]);
            const int kk = k + 2 * i;
#pragma unroll
            for (int c = 0; c < 3; ++c)
              a[u][c] = fmaf(f.x, w[c * width + kk], fmaf(f.y, w[c * width + kk + 1], a[u][c]));
          }
        }
      }
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int c = 0; c < 3; ++c) a[u][c] += __shfl_xor_sync(0xffffffffu, a[u][c], s);
    }
    // lanes 0..11: (row u = lane / 3, channel c = lane % 3)
    if (lane < 12) {
      const int u = lane / 3, c = lane - 3 * u;
      const int row = row_base + u * wstride;
      if (row < n) {
        float acc = 0.f;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu)
#pragma unroll
          for (int cc = 0; cc < 3; ++cc)
            if (uu == u && cc == c) acc = a[uu][cc];
        const float z = acc + b[c];
        const size_t idx = 3 * (size_t)row + c;
        const float yh = 1.0f / (1.0f + expf(-z));
        const float m = mask ? mask[row] : 1.0f;
        const float wgt = m + (1.0f - m) * 0.3f;
        const float d = (yh - target[idx]) * wgt;
        lsum += d * d;
        const float gi = 2.0f * d * wgt * inv_count * yh * (1.0f - yh);
        if (logits) logits[idx] = z;
        g[idx] = gi;
        lmax = fmaxf(lmax, fabsf(gi));
      }
    }
  }
  __shared__ float ssum[8], smax[8];
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    lsum += __shfl_xor_sync(0xffffffffu, lsum, s);
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, s));
  }
  if (lane == 0) {
    ssum[wib] = lsum;
    smax[wib] = lmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0.f, y = 0.f;
    for (int i = 0; i < 8; ++i) {
      x += ssum[i];
      y = fmaxf(y, smax[i]);
    }
    atomicAdd(loss, x * inv_count);
    atomicMax(amax_bits, __float_as_uint(y));
  }
}

__global__ void __launch_bounds__(256) npp_head_loss_kernel(const __half* __restrict__ hp, int ld, int width, int n,
                                                            const float* __restrict__ w, const float* __restrict__ b,
                                                            const float* __restrict__ target,
                                                            const float* __restrict__ mask, float inv_count,
                                                            float* __restrict__ logits, float* __restrict__ g,
                                                            float* __restrict__ loss, unsigned int* __restrict__ amax_bits,
                                                            const int* __restrict__ step) {
  if (step != nullptr) {   // re-launched step graph: batch *step of [iters, n, 3] targets / [iters, n] masks / [iters] losses
    const int sidx = *step;
    target += (size_t)sidx * n * 3;
    if (mask != nullptr) mask += (size_t)sidx * n;
    loss += sidx;
  }
  npp_head_loss_part(hp, ld, width, n, w, b, target, mask, inv_count, logits, g, loss, amax_bits);
}

// Last node of the re-launched step graph.
__global__ void npp_step_advance_kernel(int* step) { *step += 1; }

// max |g| of an externally supplied gradient (autograd path).
__global__ void __launch_bounds__(256) npp_amax_kernel(const float* __restrict__ g, int total,
                                                       unsigned int* __restrict__ amax_bits) {
  float lmax = 0.f;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    lmax = fmaxf(lmax, fabsf(g[idx]));
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, s));
  if ((threadIdx.x & 31) == 0 && lmax > 0.f) atomicMax(amax_bits, __float_as_uint(lmax));
}

// Backward of the RGB head: delta_P = (g . W_rgb) * snake'(z_P) * scale (fp16), plus
// dW_rgb, db_rgb (unscaled fp32) and the bias gradient of the P layer (scaled column sums).
// One warp per row, lane owns 8 consecutive columns (16-byte loads/stores); per-lane partial sums are reduced
// across the block's warps in shared memory, then one atomic per column per block.
constexpr int HEAD_BWD_ROWS = 32;   // rows per block of the stand-alone kernel: ~3.5 blocks per SM at 16 k rows
// g and amax_bits may have been written earlier in the SAME kernel (fused variant): they are read through L2
// (ld.global.cg), never through the non-coherent path.
__device__ __forceinline__ void npp_head_bwd_part(const float* g, const __half* __restrict__ hp,
                                                  const __half* __restrict__ dp, int ld, int width, int n,
                                                  const float* __restrict__ w, const unsigned int* amax_bits,
                                                  __half* __restrict__ delta, int ldd, float* head_acc /*[3*width+3]*/,
                                                  float* bias_acc /*[width]*/, int rows_per_block) {
  __shared__ float red[8][4 * 256 + 4];
  const float scale = npp_grad_scale(__uint_as_float(__ldcg(amax_bits)));
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int row_begin = blockIdx.x * rows_per_block;
  const int row_end = min(row_begin + rows_per_block, n);
  float gsum[3] = {0.f, 0.f, 0.f};
  for (int kb = 0; kb < width; kb += 256) {   // one pass for width <= 256; the barriers below are block-wide
    const int k0 = kb + lane * 8;
    const bool active = k0 < width;           // width is a multiple of 8: a lane's eight columns are all in or all out
    float wr[3][8], aw[3][8], ab[8];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        wr[c][i] = active ? w[c * width + k0 + i] : 0.f;
        aw[c][i] = 0.f;
      }
#pragma unroll
    for (int i = 0; i < 8; ++i) ab[i] = 0.f;
#pragma unroll 4
    for (int row = row_begin + wib; row < row_end && active; row += 8) {
      const float g0 = __ldcg(g + 3 * (size_t)row), g1 = __ldcg(g + 3 * (size_t)row + 1),
                  g2 = __ldcg(g + 3 * (size_t)row + 2);
      const uint4 hraw = *reinterpret_cast<const uint4*>(hp + (size_t)row * ld + k0);
      const uint4 draw = *reinterpret_cast<const uint4*>(dp + (size_t)row * ld + k0);
      const __half2* h2 = reinterpret_cast<const __half2*>(&hraw);
      const __half2* d2 = reinterpret_cast<const __half2*>(&draw);
      uint4 outv;
      __half2* o2 = reinterpret_cast<__half2*>(&outv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 h = __half22float2(h2[i]);
        const float2 d = __half22float2(d2[i]);
        const float dax = fmaf(g0, wr[0][2 * i], fmaf(g1, wr[1][2 * i], g2 * wr[2][2 * i]));
        const float day = fmaf(g0, wr[0][2 * i + 1], fmaf(g1, wr[1][2 * i + 1], g2 * wr[2][2 * i + 1]));
        const __half2 dh = __floats2half2_rn(dax * d.x * scale, day * d.y * scale);
        o2[i] = dh;
        const float2 df = __half22float2(dh);
        ab[2 * i] += df.x;
        ab[2 * i + 1] += df.y;
        aw[0][2 * i] = fmaf(g0, h.x, aw[0][2 * i]);
        aw[0][2 * i + 1] = fmaf(g0, h.y, aw[0][2 * i + 1]);
        aw[1][2 * i] = fmaf(g1, h.x, aw[1][2 * i]);
        aw[1][2 * i + 1] = fmaf(g1, h.y, aw[1][2 * i + 1]);
        aw[2][2 * i] = fmaf(g2, h.x, aw[2][2 * i]);
        aw[2][2 * i + 1] = fmaf(g2, h.y, aw[2][2 * i + 1]);
      }
      *reinterpret_cast<uint4*>(delta + (size_t)row * ldd + k0) = outv;
      if (kb == 0 && lane == 0) {
        gsum[0] += g0;
        gsum[1] += g1;
        gsum[2] += g2;
      }
    }
    // block reduction over the 8 warps (column k0+i of accumulator a lives at red[warp][a*256 + ...])
    if (kb == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        red[wib][0 * 256 + lane * 8 + i] = aw[0][i];
        red[wib][1 * 256 + lane * 8 + i] = aw[1][i];
        red[wib][2 * 256 + lane * 8 + i] = aw[2][i];
        red[wib][3 * 256 + lane * 8 + i] = ab[i];
      }
      if (lane == 0) {
        red[wib][1024] = gsum[0];
        red[wib][1025] = gsum[1];
        red[wib][1026] = gsum[2];
      }
      __syncthreads();
      for (int idx = threadIdx.x; idx < 1027; idx += blockDim.x) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][idx];
        const int a = idx >> 8, col = idx & 255;
        if (idx >= 1024) atomicAdd(head_acc + 3 * width + (idx - 1024), t);
        else if (col < width) {
          if (a < 3) atomicAdd(head_acc + a * width + col, t);
          else if (bias_acc != nullptr) atomicAdd(bias_acc + col, t);
        }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(256) npp_head_bwd_kernel(const float* __restrict__ g, const __half* __restrict__ hp,
                                                           const __half* __restrict__ dp, int ld, int width, int n,
                                                           const float* __restrict__ w,
                                                           const unsigned int* __restrict__ amax_bits,
                                                           __half* __restrict__ delta, int ldd,
                                                           float* __restrict__ head_acc, float* __restrict__ bias_acc) {
  npp_head_bwd_part(g, hp, dp, ld, width, n, w, amax_bits, delta, ldd, head_acc, bias_acc, HEAD_BWD_ROWS);
}

// Self-resetting grid barrier (sense reversal on a generation word): state[0] = arrivals, state[1] = generation.
// Needs every block of the grid to be resident (cooperative launch).
__device__ __forceinline__ void npp_grid_barrier(unsigned int* state) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int gen;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(state + 1) : "memory");
    __threadfence();
    const unsigned int prev = atomicAdd(state, 1u);
    if (prev == gridDim.x - 1) {
      state[0] = 0u;  // nobody touches the counter again before the generation moves
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(state + 1), "r"(gen + 1u) : "memory");
    } else {
      unsigned int now;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(state + 1) : "memory");
      } while (now == gen);
    }
    __threadfence();
  }
  __syncthreads();
}

// Train-step head in ONE cooperative launch: RGB head + sigmoid + masked MSE (writes g, loss, max|g|), a grid
// barrier so that every block sees the final max|g| (the fp16 delta scale), then the head backward.  The second
// read of h_P comes from L2 (8 MB at 16 k rows).
__global__ void __launch_bounds__(256, 2) npp_head_fused_kernel(const __half* __restrict__ hp, const __half* __restrict__ dp,
                                                                int ld, int width, int n, const float* __restrict__ w,
                                                                const float* __restrict__ b,
                                                                const float* __restrict__ target,
                                                                const float* __restrict__ mask, float inv_count,
                                                                float* logits, float* g, float* loss,
                                                                unsigned int* amax_bits, __half* __restrict__ delta,
                                                                int ldd, float* head_acc, float* bias_acc,
                                                                unsigned int* barrier_state) {
  npp_head_loss_part(hp, ld, width, n, w, b, target, mask, inv_count, logits, g, loss, amax_bits);
  npp_grid_barrier(barrier_state);
  const int rows_per_block = (n + (int)gridDim.x - 1) / (int)gridDim.x;
  npp_head_bwd_part(g, hp, dp, ld, width, n, w, amax_bits, delta, ldd, head_acc, bias_acc, rows_per_block);
}

// Register-resident variant for width == 256 and at most 8 * HEAD_MAXR rows per block (16 k rows on 2 x 148 blocks):
// every global load of a phase is issued up front, h_P is read once (logits AND dW_rgb come from the same
// registers), and the snake-derivative rows are fetched before the grid barrier so that their latency hides behind it.
constexpr int HEAD_MAXR = 8;
__global__ void __launch_bounds__(256, 2) npp_head_fused_reg_kernel(
    const __half* __restrict__ hp, const __half* __restrict__ dp, int ld, int n, const float* __restrict__ w,
    const float* __restrict__ b, const float* __restrict__ target, const float* __restrict__ mask, float inv_count,
    float* logits, float* g, float* loss, unsigned int* amax_bits, __half* __restrict__ delta, int ldd,
    float* head_acc /*[3*256+3]*/, float* bias_acc /*[256]*/, unsigned int* barrier_state, int rows_per_block) {
  constexpr int W = 256;
  __shared__ float red[8][4 * W + 4];
  __shared__ float ssum[8], smax[8], gs[3];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int row_begin = blockIdx.x * rows_per_block;
  const int row_end = min(row_begin + rows_per_block, n);
  if (threadIdx.x < 3) gs[threadIdx.x] = 0.f;

  // ---- phase A: logits, loss, g, dW_rgb
  uint4 hraw[HEAD_MAXR];
#pragma unroll
  for (int i = 0; i < HEAD_MAXR; ++i) {
    const int row = row_begin + wib + 8 * i;
    hraw[i] = make_uint4(0, 0, 0, 0);
    if (row < row_end) hraw[i] = *reinterpret_cast<const uint4*>(hp + (size_t)row * ld + lane * 8);
  }
  float wr[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + c * W + lane * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + c * W + lane * 8 + 4));
    wr[c][0] = w0.x; wr[c][1] = w0.y; wr[c][2] = w0.z; wr[c][3] = w0.w;
    wr[c][4] = w1.x; wr[c][5] = w1.y; wr[c][6] = w1.z; wr[c][7] = w1.w;
  }
  // lane L < 24 finishes (row slot i = L / 3, channel c = L % 3): its target / mask loads go out now as well
  const int my_i = lane / 3, my_c = lane - 3 * my_i;
  const int my_row = row_begin + wib + 8 * my_i;
  const bool mine = lane < 3 * HEAD_MAXR && my_row < row_end;
  float my_t = 0.f, my_m = 1.0f, my_b = 0.f;
  if (mine) {
    my_t = target[3 * (size_t)my_row + my_c];
    if (mask) my_m = mask[my_row];
    my_b = b[my_c];
  }
  float a[HEAD_MAXR][3];
#pragma unroll
  for (int i = 0; i < HEAD_MAXR; ++i) {
    a[i][0] = a[i][1] = a[i][2] = 0.f;
    const __half2* h2 = reinterpret_cast<const __half2*>(&hraw[i]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(h2[k]);
#pragma unroll
      for (int c = 0; c < 3; ++c) a[i][c] = fmaf(f.x, wr[c][2 * k], fmaf(f.y, wr[c][2 * k + 1], a[i][c]));
    }
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1)
#pragma unroll
    for (int i = 0; i < HEAD_MAXR; ++i)
#pragma unroll
      for (int c = 0; c < 3; ++c) a[i][c] += __shfl_xor_sync(0xffffffffu, a[i][c], s);
  float gi = 0.f, lsum = 0.f;
  {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_MAXR; ++i)
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (3 * i + c == lane) acc = a[i][c];
    if (mine) {
      const float z = acc + my_b;
      const float yh = 1.0f / (1.0f + expf(-z));
      const float wgt = my_m + (1.0f - my_m) * 0.3f;
      const float d = (yh - my_t) * wgt;
      lsum = d * d;
      gi = 2.0f * d * wgt * inv_count * yh * (1.0f - yh);
      const size_t idx = 3 * (size_t)my_row + my_c;
      if (logits) logits[idx] = z;
      g[idx] = gi;
      atomicAdd(&gs[my_c], gi);  // db_rgb
    }
  }
  float lmax = fabsf(gi);
  // dW_rgb partial sums of this warp's rows (lane owns 8 columns)
  float aw[3][8];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < 8; ++k) aw[c][k] = 0.f;
  float gr[HEAD_MAXR][3];
#pragma unroll
  for (int i = 0; i < HEAD_MAXR; ++i) {
#pragma unroll
    for (int c = 0; c < 3; ++c) gr[i][c] = __shfl_sync(0xffffffffu, gi, 3 * i + c);
    const __half2* h2 = reinterpret_cast<const __half2*>(&hraw[i]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(h2[k]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        aw[c][2 * k] = fmaf(gr[i][c], f.x, aw[c][2 * k]);
        aw[c][2 * k + 1] = fmaf(gr[i][c], f.y, aw[c][2 * k + 1]);
      }
    }
  }
  // snake-derivative rows for phase B: in flight across the barrier
  uint4 draw[HEAD_MAXR];
#pragma unroll
  for (int i = 0; i < HEAD_MAXR; ++i) {
    const int row = row_begin + wib + 8 * i;
    draw[i] = make_uint4(0, 0, 0, 0);
    if (row < row_end) draw[i] = *reinterpret_cast<const uint4*>(dp + (size_t)row * ld + lane * 8);
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    lsum += __shfl_xor_sync(0xffffffffu, lsum, s);
    lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, s));
  }
  if (lane == 0) {
    ssum[wib] = lsum;
    smax[wib] = lmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float x = 0.f, y = 0.f;
    for (int i = 0; i < 8; ++i) {
      x += ssum[i];
      y = fmaxf(y, smax[i]);
    }
    atomicAdd(loss, x * inv_count);
    if (y > 0.f) atomicMax(amax_bits, __float_as_uint(y));
  }
  npp_grid_barrier(barrier_state);

  // ---- phase B: delta_P = (g . W_rgb) * snake'(z_P) * scale, bias gradient of the P layer
  const float scale = npp_grad_scale(__uint_as_float(__ldcg(amax_bits)));
  float ab[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) ab[k] = 0.f;
#pragma unroll
  for (int i = 0; i < HEAD_MAXR; ++i) {
    const int row = row_begin + wib + 8 * i;
    if (row >= row_end) continue;
    const __half2* d2 = reinterpret_cast<const __half2*>(&draw[i]);
    uint4 outv;
    __half2* o2 = reinterpret_cast<__half2*>(&outv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 d = __half22float2(d2[k]);
      const float dax = fmaf(gr[i][0], wr[0][2 * k], fmaf(gr[i][1], wr[1][2 * k], gr[i][2] * wr[2][2 * k]));
      const float day = fmaf(gr[i][0], wr[0][2 * k + 1], fmaf(gr[i][1], wr[1][2 * k + 1], gr[i][2] * wr[2][2 * k + 1]));
      const __half2 dh = __floats2half2_rn(dax * d.x * scale, day * d.y * scale);
      o2[k] = dh;
      const float2 df = __half22float2(dh);
      ab[2 * k] += df.x;
      ab[2 * k + 1] += df.y;
    }
    *reinterpret_cast<uint4*>(delta + (size_t)row * ldd + lane * 8) = outv;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[wib][0 * W + lane * 8 + k] = aw[0][k];
    red[wib][1 * W + lane * 8 + k] = aw[1][k];
    red[wib][2 * W + lane * 8 + k] = aw[2][k];
    red[wib][3 * W + lane * 8 + k] = ab[k];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 4 * W; idx += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += red[q][idx];
    const int acc_i = idx >> 8, col = idx & (W - 1);
    if (acc_i < 3) atomicAdd(head_acc + acc_i * W + col, t);
    else if (bias_acc != nullptr) atomicAdd(bias_acc + col, t);
  }
  if (threadIdx.x < 3) atomicAdd(head_acc + 3 * W + threadIdx.x, gs[threadIdx.x]);
}

// Masked MSE on the network OUTPUT (after the sigmoid), forward + gradient in one pass: the 'l2' branch of
// models/mse_calculator.py:13-27 as the reference scripts call it, img2mse(pred, gt, 'l2', None, mask):
//   d = (x - y) * (m + 0.3 (1 - m));  loss = mean(d^2) over [n,3];  dL/dx = 2 d (m + 0.3 (1 - m)) / (3 n)
__global__ void __launch_bounds__(256) npp_l2_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                     const float* __restrict__ mask, int n, float inv_count,
                                                     float* __restrict__ gx, float* __restrict__ loss) {
  float lsum = 0.f;
  const int total = 3 * n;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const float m = mask ? mask[idx / 3] : 1.0f;
    const float w = m + (1.0f - m) * 0.3f;
    const float d = (x[idx] - y[idx]) * w;
    lsum += d * d;
    gx[idx] = 2.0f * d * w * inv_count;
  }
  __shared__ float ssum[8];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if ((threadIdx.x & 31) == 0) ssum[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += ssum[i];
    atomicAdd(loss, t * inv_count);
  }
}

// ------------------------------------------------------------- adaptive robust pixel loss (Barron)
// models/mse_calculator.py:24-25 with --loss_type robust_loss_adaptive (the reference default):
//   d = (x - y) * (m + 0.3 (1 - m));  loss = mean over [N,3] of  rho(d, alpha_c, s_c) + log s_c + log Z(alpha_c)
//   rho(d, a, s) = (|a-2| / a) * (((d/s)^2 / |a-2| + 1)^(a/2) - 1)            (robust_loss_pytorch/general.py:84-118)
//   alpha_c = lo + (hi - lo) sigmoid(latent_alpha_c),  s_c = (ref - lo_s) softplus(latent_scale_c + log(e - 1)) + lo_s
//                                                               (adaptive.py:138-175, util.py:64-95)
//   log Z: distribution.py:143-171; here a cubic Hermite table of the exact integral (tools/make_robust_logz_table.py).
struct RobustCfg {
  float alpha_lo, alpha_hi, scale_lo, scale_ref;
  int n_knots;
  float alpha_max;  // table covers [0, alpha_max]
};

__device__ __forceinline__ void npp_robust_channel(const float* __restrict__ latent_alpha,
                                                   const float* __restrict__ latent_scale, RobustCfg cfg,
                                                   const float* __restrict__ zval, const float* __restrict__ zder, int c,
                                                   float& alpha, float& scale, float& logz, float& dlogz,
                                                   float& dalpha_dlat, float& dscale_dlat) {
  const float la = latent_alpha[c], ls = latent_scale[c];
  const float sg = 1.0f / (1.0f + expf(-la));
  alpha = cfg.alpha_lo + (cfg.alpha_hi - cfg.alpha_lo) * sg;
  dalpha_dlat = (cfg.alpha_hi - cfg.alpha_lo) * sg * (1.0f - sg);
  const float t = ls + 0.541324854612918f;  // inv_softplus(1) = log(e - 1)
  const float sp = t > 20.0f ? t : log1pf(expf(t));
  scale = (cfg.scale_ref - cfg.scale_lo) * sp + cfg.scale_lo;
  dscale_dlat = (cfg.scale_ref - cfg.scale_lo) / (1.0f + expf(-t));
  // cubic Hermite interpolation of log Z and its derivative
  const float h = cfg.alpha_max / (float)(cfg.n_knots - 1);
  float pos = fminf(fmaxf(alpha / h, 0.0f), (float)(cfg.n_knots - 1) - 1e-3f);
  const int i = (int)pos;
  const float u = pos - (float)i;
  const float v0 = zval[i], v1 = zval[i + 1], d0 = zder[i] * h, d1 = zder[i + 1] * h;
  const float u2 = u * u, u3 = u2 * u;
  logz = (2 * u3 - 3 * u2 + 1) * v0 + (u3 - 2 * u2 + u) * d0 + (-2 * u3 + 3 * u2) * v1 + (u3 - u2) * d1;
  dlogz = ((6 * u2 - 6 * u) * v0 + (3 * u2 - 4 * u + 1) * d0 + (-6 * u2 + 6 * u) * v1 + (3 * u2 - 2 * u) * d1) / h;
}

// acc[0..2] = sum rho, acc[3..5] = sum d rho / d alpha, acc[6..8] = sum d rho / d scale (per channel)
__global__ void __launch_bounds__(256) npp_robust_loss_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                              const float* __restrict__ mask, int n,
                                                              const float* __restrict__ latent_alpha,
                                                              const float* __restrict__ latent_scale, RobustCfg cfg,
                                                              const float* __restrict__ zval,
                                                              const float* __restrict__ zder, float inv_count,
                                                              float* __restrict__ gx, float* __restrict__ acc) {
  __shared__ float s_alpha[3], s_scale[3];
  __shared__ float red[8][9];
  if (threadIdx.x < 3) {
    float a, s, lz, dlz, da, ds;
    npp_robust_channel(latent_alpha, latent_scale, cfg, zval, zder, threadIdx.x, a, s, lz, dlz, da, ds);
    s_alpha[threadIdx.x] = a;
    s_scale[threadIdx.x] = s;
  }
  __syncthreads();
  float sum[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) sum[k] = 0.f;
  // one row (three channels) per thread and iteration
  const int stride = gridDim.x * blockDim.x * 3;
  for (int base = (blockIdx.x * blockDim.x + threadIdx.x) * 3; base < 3 * n; base += stride) {
    const int row = base / 3;
    const float m = mask ? mask[row] : 1.0f;
    const float w = m + (1.0f - m) * 0.3f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = s_alpha[c], s = s_scale[c];
      const float d = (x[base + c] - y[base + c]) * w;
      const float b = fmaxf(fabsf(a - 2.0f), 1.1920929e-07f);
      const float q = (d / s) * (d / s);
      const float u = q / b + 1.0f;
      const float lu = logf(u);
      const float p = expf(0.5f * a * lu);
      const float pu = p / u;
      sum[c] += (b / a) * (p - 1.0f);
      sum[3 + c] += (-2.0f / (a * a)) * (p - 1.0f) + (b / a) * p * (0.5f * lu + 0.5f * a * (q / (b * b)) / u);
      sum[6 + c] += -(q / s) * pu;
      gx[base + c] = (d / (s * s)) * pu * w * inv_count;
    }
  }
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], o);
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int k = 0; k < 9; ++k) red[threadIdx.x >> 5][k] = sum[k];
  __syncthreads();
  if (threadIdx.x < 9) {
    float t = 0.f;
    for (int wq = 0; wq < 8; ++wq) t += red[wq][threadIdx.x];
    atomicAdd(acc + threadIdx.x, t);
  }
}

// out[0] = loss, out[1..3] = dL/d latent_alpha, out[4..6] = dL/d latent_scale
__global__ void npp_robust_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ latent_alpha,
                                           const float* __restrict__ latent_scale, RobustCfg cfg,
                                           const float* __restrict__ zval, const float* __restrict__ zder,
                                           float inv_count, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float loss = 0.f;
  for (int c = 0; c < 3; ++c) {
    float a, s, lz, dlz, da, ds;
    npp_robust_channel(latent_alpha, latent_scale, cfg, zval, zder, c, a, s, lz, dlz, da, ds);
    loss += inv_count * acc[c] + (logf(s) + lz) * (1.0f / 3.0f);
    out[1 + c] = (inv_count * acc[3 + c] + dlz * (1.0f / 3.0f)) * da;
    out[4 + c] = (inv_count * acc[6 + c] + (1.0f / 3.0f) / s) * ds;
  }
  out[0] = loss;
}

// ------------------------------------------------------------- gradients / Adam
// Sum of the split-K partials of one element in a fixed order; all loads are issued before the adds.
constexpr int NPP_MAX_SPLITS = 12;
__device__ __forceinline__ float npp_sum_splits(const float* __restrict__ p, int n_splits, long long stride) {
  float v[NPP_MAX_SPLITS];
#pragma unroll
  for (int k = 0; k < NPP_MAX_SPLITS; ++k) v[k] = k < n_splits ? __ldg(p + k * stride) : 0.f;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < NPP_MAX_SPLITS; ++k) s += v[k];
  return s;
}

struct FinalizeLayer {
  long long w_off, b_off;    // arena offsets (floats) of weight [out, in_ref] and bias [out]
  long long pg_off;          // offset of this layer's [out, kpad] block inside a partial slab
  long long bg_off;          // offset of the scaled bias-gradient accumulator
  int out, in_ref, kpad;
  int split_col;             // reference columns < split_col map to off0 + c, others to off1 + (c - split_col)
  int off0, off1;
};

// grads[w] = (sum over split-K slabs) / scale ; grads[b] = bias_acc / scale.
// grid = (row groups, layers); each block walks whole rows so the slab reads are coalesced.
__global__ void __launch_bounds__(256) npp_grad_finalize_kernel(const FinalizeLayer* __restrict__ layers,
                                                                const float* __restrict__ partial, int n_splits,
                                                                long long slab_stride, const float* __restrict__ bias_acc,
                                                                const unsigned int* __restrict__ amax_bits,
                                                                float* __restrict__ grads, float step_inv_count) {
  const FinalizeLayer L = layers[blockIdx.y];
  // step_inv_count > 0: the deltas carry the fused step's scale (npp_step_amax of the previous step's maximum)
  const float inv = 1.0f / npp_grad_scale(step_inv_count > 0.f ? npp_step_amax(amax_bits, step_inv_count)
                                                                : __uint_as_float(*amax_bits));
  for (int o = blockIdx.x; o < L.out; o += gridDim.x) {
    const float* prow = partial + L.pg_off + (long long)o * L.kpad;
    float* grow = grads + L.w_off + (long long)o * L.in_ref;
    for (int c = threadIdx.x; c < L.in_ref; c += blockDim.x) {
      const int pc = c < L.split_col ? L.off0 + c : L.off1 + (c - L.split_col);
      grow[c] = npp_sum_splits(prow + pc, n_splits, slab_stride) * inv;
    }
  }
  if (blockIdx.x == 0)
    for (int o = threadIdx.x; o < L.out; o += blockDim.x) grads[L.b_off + o] = bias_acc[L.bg_off + o] * inv;
}

// Single pass for the fused train step: split-K slabs -> gradient -> Adam -> fp32 master + fp16 shadows.
// Replaces grad_finalize + adam + shadow (3 kernels, ~2x the bytes) when nobody needs the gradient arena.
struct UpdateLayer {
  long long w_off, b_off, pg_off, bg_off;
  int out, in_ref, kpad;       // out: reference out_features (rows of the arena tensor)
  int split_col, off0, off1;
  __half* wf;
  __half* wt;
  int wt_ld;                   // row pitch of wt: out rounded up to the GEMM tile (256)
  int t_lo, t_hi, t_row0;
  int t_lo2, t_hi2, t_row02;
};
struct AdamScalars {
  float beta1, beta2, step_size, inv_sqrt_bc2, eps;
};
// Fused step (forward + head + backward in one launch, see EPI_SNAKE_HEAD): the update kernel is the last reader of
// the step's accumulators, so it clears them for the next step, hands the loss out, and rotates the max|g| ring.
struct StepReset {
  int on;                    // 0: legacy behaviour (nothing below is touched, amax_bits holds this step's maximum)
  float inv_count;
  float* loss_acc;
  float* loss_out;
  unsigned int* amax_clear;
  unsigned int* ring;        // re-launched step graph (step != nullptr): ring slots follow (seq0 + *step) % 3 and the loss
  int seq0;                  //   goes to loss_out[*step]
  int by_step;
};
__device__ __forceinline__ float npp_adam1(float p, float g, float& m, float& v, const AdamScalars a) {
  m = m + (g - m) * (1.0f - a.beta1);
  v = v * a.beta2 + (1.0f - a.beta2) * g * g;
  return p - a.step_size * (m / (sqrtf(v) * a.inv_sqrt_bc2 + a.eps));
}
// All layers of one plan, passed by value (constant bank: no dependent global load before the first slab read).
constexpr int NPP_MAX_UPDATE_LAYERS = 24;
struct UpdateTable {
  UpdateLayer L[NPP_MAX_UPDATE_LAYERS];
  int tile_begin[NPP_MAX_UPDATE_LAYERS + 1];  // prefix sums of 32 x 128 tiles; block b works on exactly one tile
  int n_layers;
};

// One block = one tile of 32 weight rows x 128 PADDED columns (blockIdx.x == total tiles: the rgb_linear head).
// A thread owns 4 consecutive padded columns of one row per pass, so the split-K slabs are read with aligned
// float4 loads; padded columns map back to reference columns per segment (segment boundaries are multiples of 64,
// so the four columns never straddle one).  The four passes run as two groups of two: every load of a group
// (S slab float4 + master/m/v) is issued before the first result is consumed, which is what keeps enough bytes in
// flight to approach HBM bandwidth.
template <int S>
__global__ void __launch_bounds__(256) npp_fused_update_kernel(const __grid_constant__ UpdateTable tab,
                                                               const float* __restrict__ partial, long long slab_stride,
                                                               float* bias_acc, float* head_acc, long long rgb_w_off,
                                                               long long rgb_b_off, int head_width,
                                                               const unsigned int* __restrict__ amax_bits,
                                                               float* __restrict__ params, float* __restrict__ grads,
                                                               float* __restrict__ m, float* __restrict__ v,
                                                               AdamScalars ad_arg,
                                                               const AdamScalars* __restrict__ ad_table,
                                                               const int* __restrict__ step, StepReset rs) {
  __shared__ float tile[32][129];
  grid_launch_dependents();   // programmatic dependent launch: the next kernel's prologue may overlap this kernel
  grid_dependency_wait();
  // re-launched step graph: the scalars of Adam step *step come from a table filled by npp_fit_run
  const AdamScalars ad = ad_table != nullptr ? ad_table[*step] : ad_arg;
  const int total_tiles = tab.tile_begin[tab.n_layers];
  if ((int)blockIdx.x >= total_tiles) {  // rgb_linear: unscaled fp32 accumulators written by the head backward
    const int total = 3 * head_width + 3;
    float* hacc = head_acc;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const long long idx = i < 3 * head_width ? rgb_w_off + i : rgb_b_off + (i - 3 * head_width);
      const float g = head_acc[i];
      if (rs.on) hacc[i] = 0.f;   // the consumer clears the step accumulators (nobody else reads this entry)
      if (grads) grads[idx] = g;
      params[idx] = npp_adam1(params[idx], g, m[idx], v[idx], ad);
    }
    if (rs.on && threadIdx.x == 0) {
      float* lout = rs.loss_out;
      unsigned int* clr = rs.amax_clear;
      if (rs.by_step) {
        const int sidx = *step;
        if (lout != nullptr) lout += sidx;
        clr = rs.ring + (rs.seq0 + sidx + 1) % 3;
      }
      if (lout != nullptr) *lout = *rs.loss_acc;
      *rs.loss_acc = 0.f;
      *clr = 0u;                  // ring slot the step after the next one accumulates into
    }
    return;
  }
  int li = 0;
  while ((int)blockIdx.x >= tab.tile_begin[li + 1]) ++li;
  const UpdateLayer& L = tab.L[li];
  const int t = (int)blockIdx.x - tab.tile_begin[li];
  // fused step: the deltas were scaled with the PREVIOUS step's maximum (npp_step_amax), else with this step's
  const unsigned int* prev_bits = (rs.on && rs.by_step) ? rs.ring + (rs.seq0 + *step + 2) % 3 : amax_bits;
  const float inv = 1.0f / npp_grad_scale(rs.on ? npp_step_amax(prev_bits, rs.inv_count) : __uint_as_float(*amax_bits));
  const int tiles_c = (L.kpad + 127) >> 7;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 column groups x 8 rows per pass
  const bool two_seg = L.in_ref > L.split_col;
  const int r0 = (t / tiles_c) << 5, pc0 = (t % tiles_c) << 7;
  const int pc = pc0 + 4 * tx;
  // this thread's four padded columns -> reference columns [c, c+4) clipped to lim (-1: padding only)
  int c = -1, lim = 0;
  if (pc < L.kpad) {
    if (two_seg && pc >= L.off1) {
      c = pc - L.off1 + L.split_col;
      lim = L.in_ref;
    } else {
      c = pc - L.off0;
      lim = L.split_col;
    }
    if (c >= lim) c = -1;
  }
  // float2 access to the fp32 arenas needs even offsets (true for every reference shape: widths are even)
  const bool vec2 = ((L.in_ref | L.split_col | (int)(L.w_off & 1)) & 1) == 0;
  const long long s4 = slab_stride >> 2;
#pragma unroll
  for (int grp = 0; grp < 2; ++grp) {
    float4 acc[2][S];
    float pv[2][4], mv[2][4], vv[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = r0 + ty + 8 * (2 * grp + u);
#pragma unroll
      for (int j = 0; j < 4; ++j) pv[u][j] = mv[u][j] = vv[u][j] = 0.f;
      if (c >= 0) {
        const float4* pp = reinterpret_cast<const float4*>(partial + L.pg_off + (long long)r * L.kpad + pc);
#pragma unroll
        for (int k = 0; k < S; ++k) acc[u][k] = __ldg(pp + k * s4);
        const long long base = L.w_off + (long long)r * L.in_ref + c;
        if (vec2) {
#pragma unroll
          for (int j = 0; j < 4; j += 2)
            if (c + j < lim) {
              const float2 a = *reinterpret_cast<const float2*>(params + base + j);
              const float2 b = *reinterpret_cast<const float2*>(m + base + j);
              const float2 d = *reinterpret_cast<const float2*>(v + base + j);
              pv[u][j] = a.x; pv[u][j + 1] = a.y;
              mv[u][j] = b.x; mv[u][j + 1] = b.y;
              vv[u][j] = d.x; vv[u][j + 1] = d.y;
            }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < lim) {
              pv[u][j] = params[base + j];
              mv[u][j] = m[base + j];
              vv[u][j] = v[base + j];
            }
        }
      } else {
#pragma unroll
        for (int k = 0; k < S; ++k) acc[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int r = r0 + ty + 8 * (2 * grp + u);
      float pn[4] = {0.f, 0.f, 0.f, 0.f};
      if (c >= 0) {
        float g4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < S; ++k) {  // fixed order: deterministic
          g4[0] += acc[u][k].x;
          g4[1] += acc[u][k].y;
          g4[2] += acc[u][k].z;
          g4[3] += acc[u][k].w;
        }
        const long long base = L.w_off + (long long)r * L.in_ref + c;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          g4[j] *= inv;
          if (c + j < lim) pn[j] = npp_adam1(pv[u][j], g4[j], mv[u][j], vv[u][j], ad);
        }
        if (vec2) {
#pragma unroll
          for (int j = 0; j < 4; j += 2)
            if (c + j < lim) {
              if (grads) *reinterpret_cast<float2*>(grads + base + j) = make_float2(g4[j], g4[j + 1]);
              *reinterpret_cast<float2*>(params + base + j) = make_float2(pn[j], pn[j + 1]);
              *reinterpret_cast<float2*>(m + base + j) = make_float2(mv[u][j], mv[u][j + 1]);
              *reinterpret_cast<float2*>(v + base + j) = make_float2(vv[u][j], vv[u][j + 1]);
            }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c + j < lim) {
              if (grads) grads[base + j] = g4[j];
              params[base + j] = pn[j];
              m[base + j] = mv[u][j];
              v[base + j] = vv[u][j];
            }
        }
      }
      if (pc < L.kpad) {
        // forward shadow: four halves (padding columns are written as the zeros they must stay)
        const __half2 lo = __floats2half2_rn(pn[0], pn[1]), hi = __floats2half2_rn(pn[2], pn[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<const unsigned int*>(&lo);
        pk.y = *reinterpret_cast<const unsigned int*>(&hi);
        *reinterpret_cast<uint2*>(L.wf + (long long)r * L.kpad + pc) = pk;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) tile[ty + 8 * (2 * grp + u)][4 * tx + j] = pn[j];
    }
  }
  __syncthreads();
  if (L.wt != nullptr) {
    // transposed shadow for dgrad: row = reference input column (within the ranges that need a gradient); a
    // half-warp writes one column as 16 half2 (two weight rows each), so a warp covers two columns per pass
    const int rp = 2 * (tx & 15);
#pragma unroll 4
    for (int i = 0; i < 8; ++i) {
      const int pcl = 2 * ty + (tx >> 4) + 16 * i;
      const int pcc = pc0 + pcl;
      if (pcc >= L.kpad) continue;
      int cc, ll;
      if (two_seg && pcc >= L.off1) {
        cc = pcc - L.off1 + L.split_col;
        ll = L.in_ref;
      } else {
        cc = pcc - L.off0;
        ll = L.split_col;
      }
      if (cc >= ll) continue;
      int trow = -1;
      if (cc >= L.t_lo && cc < L.t_hi) trow = L.t_row0 + (cc - L.t_lo);
      else if (cc >= L.t_lo2 && cc < L.t_hi2) trow = L.t_row02 + (cc - L.t_lo2);
      if (trow >= 0)
        *reinterpret_cast<__half2*>(L.wt + (long long)trow * L.wt_ld + r0 + rp) =
            __floats2half2_rn(tile[rp][pcl], tile[rp + 1][pcl]);
    }
  }
  if (t == 0) {
    float* bacc = bias_acc;
    for (int o = threadIdx.x; o < L.out; o += blockDim.x) {
      const long long idx = L.b_off + o;
      const float g = bias_acc[L.bg_off + o] * inv;
      if (rs.on) bacc[L.bg_off + o] = 0.f;
      if (grads) grads[idx] = g;
      params[idx] = npp_adam1(params[idx], g, m[idx], v[idx], ad);
    }
  }
}

// End of a three-phase (data-parallel) step: hand the loss out, clear the step accumulators, rotate the max|g| ring.
__global__ void npp_step_finish_kernel(float* acc, int acc_n, float* loss_acc, float* loss_out, unsigned int* amax_clear) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < acc_n; i += gridDim.x * blockDim.x) acc[i] = 0.f;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (loss_out != nullptr) *loss_out = *loss_acc;
    *loss_acc = 0.f;
    *amax_clear = 0u;
  }
}

__global__ void npp_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// torch.optim.Adam single-tensor arithmetic (torch/optim/adam.py _single_tensor_adam, constructed at
// reference models/helpers.py:164): m.lerp_(g, 1-b1); v = b2 v + (1-b2) g g;
// p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) npp_adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ m, float* __restrict__ v, long long n,
                                                       float beta1, float beta2, float step_size, float inv_sqrt_bc2,
                                                       float eps, int vec4) {
  const long long n4 = vec4 ? n >> 2 : 0;   // vec4 == 0: some pointer is not 16-byte aligned, everything takes the scalar loop
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
#define NPP_ADAM1(c)                                                   \
  mm.c = mm.c + (gg.c - mm.c) * (1.0f - beta1);                        \
  vv.c = vv.c * beta2 + (1.0f - beta2) * gg.c * gg.c;                  \
  pp.c = pp.c - step_size * (mm.c / (sqrtf(vv.c) * inv_sqrt_bc2 + eps));
    NPP_ADAM1(x) NPP_ADAM1(y) NPP_ADAM1(z) NPP_ADAM1(w)
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  // tail (arena sizes are padded to 4, kept for safety)
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float mm = m[i] + (g[i] - m[i]) * (1.0f - beta1);
    float vv = v[i] * beta2 + (1.0f - beta2) * g[i] * g[i];
    p[i] -= step_size * (mm / (sqrtf(vv) * inv_sqrt_bc2 + eps));
    m[i] = mm;
    v[i] = vv;
  }
#undef NPP_ADAM1
}

// fp32 master weights -> fp16 shadows: K-padded forward copy [out, kpad] and the transposed
// copy [in_sub, out] that the dgrad GEMMs read K-major.
struct ShadowLayer {
  long long w_off;
  int out, in_ref, kpad;
  int split_col, off0, off1;  // reference column -> padded column (same rule as FinalizeLayer)
  __half* wf;                 // [out, kpad]
  __half* wt;                 // [wt_rows, wt_ld] or null
  int wt_ld;                  // out rounded up to the GEMM tile (256); columns >= out stay zero
  int t_lo, t_hi, t_row0;     // reference columns [t_lo, t_hi) go to wt rows t_row0 + (c - t_lo)
  int t_lo2, t_hi2, t_row02;  // optional second range
};

__global__ void __launch_bounds__(256) npp_shadow_kernel(const ShadowLayer* __restrict__ layers,
                                                         const float* __restrict__ params) {
  __shared__ float tile[32][33];
  const ShadowLayer L = layers[blockIdx.y];
  const int tiles_c = (L.in_ref + 31) / 32;
  const int tiles_r = L.out / 32;
  for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
    const int r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r0 + ty + 8 * i, c = c0 + tx;
      float v = 0.f;
      if (c < L.in_ref) {
        v = params[L.w_off + (long long)r * L.in_ref + c];
        const int pc = c < L.split_col ? L.off0 + c : L.off1 + (c - L.split_col);
        L.wf[(long long)r * L.kpad + pc] = __float2half_rn(v);
      }
      tile[ty + 8 * i][tx] = v;
    }
    __syncthreads();
    if (L.wt != nullptr) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + 8 * i, r = r0 + tx;
        int trow = -1;
        if (c >= L.t_lo && c < L.t_hi) trow = L.t_row0 + (c - L.t_lo);
        else if (c >= L.t_lo2 && c < L.t_hi2) trow = L.t_row02 + (c - L.t_lo2);
        if (trow >= 0) L.wt[(long long)trow * L.wt_ld + r] = __float2half_rn(tile[tx][ty + 8 * i]);
      }
    }
  }
}

}  // namespace npp
