// mFormerV0 (RelativeAttention variant) inference kernels that the V1 path does not
// already provide.  NHWC activations throughout.
//   lnx_im2col3x3        dense 3x3 convs (stem, overlap patch embed) as gather + tensor-core GEMM
//   lnx_maxpool3s2       MaxPool2d(3, 2, 1)
//   lnx_dwconv3_fwd      depthwise 3x3 (stride 1 / 2, TF "same" static padding) + folded BN + swish, with the
//                        squeeze-excite global average pool accumulated in the same pass
//   lnx_se_scale         x * sigmoid(gate[b, c])
//   lnx_attn_bias_fwd    softmax(scale q k^T + bias[h]) v for any head_dim <= 128, reading q/k/v straight out of the
//                        [B, N, 3, heads, hd] qkv GEMM output (no split / transpose pass)
// All HBM bound except the attention, which is a CUDA-core kernel in this round (head_dim 48 / 96).
#include <stdlib.h>

#include "lnx_common.cuh"

using namespace lnx;

namespace {

template <typename T>
__device__ __forceinline__ void zero16(Vec16<T>& v) {
#pragma unroll
  for (int i = 0; i < Vec16<T>::N; ++i) v.set(i, 0.f);
}

// ---------------------------------------------------------------- im2col 3x3 (pad 1), NHWC, C % V == 0
template <typename T>
__global__ void im2col3_nhwc_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int H, int W, int C, int stride, int Ho, int Wo,
                                    int Kpad) {
  constexpr int V = Vec16<T>::N;
  const int cv = C / V;
  const long long total = (long long)B * Ho * Wo * 9 * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long r = i / cv;
    const int tap = (int)(r % 9);
    r /= 9;  // output pixel index
    const int wo = (int)(r % Wo);
    const int ho = (int)((r / Wo) % Ho);
    const long long b = r / ((long long)Wo * Ho);
    const int h = ho * stride - 1 + tap / 3, w = wo * stride - 1 + tap % 3;
    Vec16<T> v;
    if (h >= 0 && h < H && w >= 0 && w < W) v = ld16(x + ((b * H + h) * W + w) * C + c * V);
    else zero16(v);
    st16(out + r * Kpad + tap * C + c * V, v);
  }
}

// odd shapes (C not a multiple of the vector width, or zero-padded K): one thread per output element
template <typename T>
__global__ void im2col3_nhwc_scalar_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int H, int W, int C, int stride, int Ho, int Wo,
                                           int Kpad) {
  const long long total = (long long)B * Ho * Wo * Kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % Kpad);
    const long long r = i / Kpad;
    T v = from_f32<T>(0.f);
    if (col < 9 * C) {
      const int tap = col / C, c = col % C;
      const int wo = (int)(r % Wo);
      const int ho = (int)((r / Wo) % Ho);
      const long long b = r / ((long long)Wo * Ho);
      const int h = ho * stride - 1 + tap / 3, w = wo * stride - 1 + tap % 3;
      if (h >= 0 && h < H && w >= 0 && w < W) v = x[((b * H + h) * W + w) * C + c];
    }
    out[i] = v;
  }
}

// first conv: fp32 NCHW image, tiny Cin; one thread per output row, columns (kh, kw, c), zero padded to Kpad
template <typename T>
__global__ void im2col3_nchw_kernel(const float* __restrict__ x, T* __restrict__ out, int B, int H, int W, int C, int stride, int Ho, int Wo,
                                    int Kpad) {
  const long long total = (long long)B * Ho * Wo;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(r % Wo);
    const int ho = (int)((r / Wo) % Ho);
    const long long b = r / ((long long)Wo * Ho);
    T* dst = out + r * Kpad;
    int col = 0;
    for (int tap = 0; tap < 9; ++tap) {
      const int h = ho * stride - 1 + tap / 3, w = wo * stride - 1 + tap % 3;
      const bool ok = h >= 0 && h < H && w >= 0 && w < W;
      for (int c = 0; c < C; ++c, ++col) dst[col] = from_f32<T>(ok ? x[((b * C + c) * H + h) * W + w] : 0.f);
    }
    for (; col < Kpad; ++col) dst[col] = from_f32<T>(0.f);
  }
}

// RGB image (C = 3, Kpad = 32, bf16 rows of 64 bytes): the 27 taps are gathered into registers (adjacent threads read
// adjacent pixels of the same image row) and the row leaves as four 16-byte stores instead of 32 two-byte ones
__global__ void __launch_bounds__(256) im2col3_rgb_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ out, int B, int H, int W, int stride,
                                                               int Ho, int Wo) {
  const long long total = (long long)B * Ho * Wo;
  for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < total; r += (long long)gridDim.x * blockDim.x) {
    const int wo = (int)(r % Wo);
    const int ho = (int)((r / Wo) % Ho);
    const long long b = r / ((long long)Wo * Ho);
    const float* img = x + b * 3 * (long long)H * W;
    float v[32];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int h = ho * stride - 1 + tap / 3, w = wo * stride - 1 + tap % 3;
      const bool ok = h >= 0 && h < H && w >= 0 && w < W;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[tap * 3 + c] = ok ? __ldg(img + ((long long)c * H + h) * W + w) : 0.f;
    }
#pragma unroll
    for (int i = 27; i < 32; ++i) v[i] = 0.f;
    uint4* dst = reinterpret_cast<uint4*>(out + r * 32);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      Vec16<bf16> o;
#pragma unroll
      for (int e = 0; e < 8; ++e) o.set(e, v[q * 8 + e]);
      st16(reinterpret_cast<bf16*>(dst + q), o);
    }
  }
}

// ---------------------------------------------------------------- maxpool 3x3 s2 p1
template <typename T>
__global__ void maxpool3s2_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int C, int Ho, int Wo) {
  constexpr int V = Vec16<T>::N;
  const int cv = C / V;
  const long long total = (long long)B * Ho * Wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cv);
    long long r = i / cv;
    const int wo = (int)(r % Wo);
    const int ho = (int)((r / Wo) % Ho);
    const long long b = r / ((long long)Wo * Ho);
    float m[V];
#pragma unroll
    for (int e = 0; e < V; ++e) m[e] = -INFINITY;
    for (int dh = 0; dh < 3; ++dh) {
      const int h = ho * 2 - 1 + dh;
      if (h < 0 || h >= H) continue;
      for (int dw = 0; dw < 3; ++dw) {
        const int w = wo * 2 - 1 + dw;
        if (w < 0 || w >= W) continue;
        const Vec16<T> v = ld16(x + ((b * H + h) * W + w) * C + c * V);
#pragma unroll
        for (int e = 0; e < V; ++e) m[e] = fmaxf(m[e], v.get(e));
      }
    }
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < V; ++e) o.set(e, m[e]);
    st16(y + r * C + c * V, o);
  }
}

// ---------------------------------------------------------------- depthwise 3x3 + affine + swish (+ SE pool sums)
// grid (row groups, B); block (C / V, ROWS): a thread owns one channel vector of one output row and slides a 3 x 3
// window of raw input vectors along it (3 * stride new 16-byte loads per output instead of 9); the per-row partial
// sums of the squeeze-excite average pool are folded across the block's rows in shared memory, one atomic per
// (block, channel).
template <typename T, int STRIDE>
__global__ void dwconv3_kernel(const T* __restrict__ x, const float* __restrict__ w9c, const float* __restrict__ scale,
                               const float* __restrict__ shift, T* __restrict__ y, float* __restrict__ pool_sum, int H, int W, int C, int pad_t,
                               int pad_l, int Ho, int Wo, int act) {
  constexpr int V = Vec16<T>::N;
  extern __shared__ float s_pool[];  // [blockDim.y][C] when pool_sum
  const int c0 = threadIdx.x * V;
  const long long b = blockIdx.y;
  const int ho = blockIdx.x * blockDim.y + threadIdx.y;
  float wr[9][V], sc[V], sh[V], psum[V];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int e = 0; e < V; ++e) wr[t][e] = w9c[t * C + c0 + e];
#pragma unroll
  for (int e = 0; e < V; ++e) {
    sc[e] = scale ? scale[c0 + e] : 1.f;
    sh[e] = shift ? shift[c0 + e] : 0.f;
    psum[e] = 0.f;
  }
  if (ho < Ho) {
    const T* xb = x + b * (long long)H * W * C + c0;
    auto load_col = [&](int w, Vec16<T>* col) {  // the three input rows of this output row at input column w (zero outside)
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        const int h = ho * STRIDE - pad_t + dh;
        if (h >= 0 && h < H && w >= 0 && w < W) col[dh] = ld16(xb + ((long long)h * W + w) * C);
        else zero16(col[dh]);
      }
    };
    // win = the three input columns of the current output; pre = the column(s) the NEXT output adds, loaded one
    // iteration ahead so the global-load latency overlaps the 72 FMAs + swish of the current pixel
    Vec16<T> win[3][3], pre[STRIDE][3];  // [column][row]
    load_col(-pad_l, win[0]);
    load_col(1 - pad_l, win[1]);
    load_col(2 - pad_l, win[2]);
    for (int wo = 0; wo < Wo; ++wo) {
      const int w0 = wo * STRIDE - pad_l;
      if (wo + 1 < Wo) {
#pragma unroll
        for (int k = 0; k < STRIDE; ++k) load_col(w0 + 3 + k, pre[k]);
      }
      float acc[V];
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll
      for (int dw = 0; dw < 3; ++dw)
#pragma unroll
        for (int dh = 0; dh < 3; ++dh)
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = fmaf(win[dw][dh].get(e), wr[dh * 3 + dw][e], acc[e]);
      Vec16<T> o;
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float v = fmaf(acc[e], sc[e], sh[e]);
        if (act == 1) v = v / (1.f + (sizeof(T) == 4 ? expf(-v) : __expf(-v)));
        o.set(e, v);
        psum[e] += o.get(e);  // pool what the next layer will read (the rounded value)
      }
      st16(y + ((b * Ho + ho) * (long long)Wo + wo) * C + c0, o);
#pragma unroll
      for (int dh = 0; dh < 3; ++dh) {
        if (STRIDE == 1) {
          win[0][dh] = win[1][dh];
          win[1][dh] = win[2][dh];
          win[2][dh] = pre[0][dh];
        } else {
          win[0][dh] = win[2][dh];
          win[1][dh] = pre[0][dh];
          win[2][dh] = pre[STRIDE - 1][dh];
        }
      }
    }
  }
  if (pool_sum) {
#pragma unroll
    for (int e = 0; e < V; ++e) s_pool[threadIdx.y * C + c0 + e] = psum[e];
    __syncthreads();
    if (threadIdx.y == 0) {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float t = 0.f;
        for (int r = 0; r < (int)blockDim.y; ++r) t += s_pool[r * C + c0 + e];
        atomicAdd(pool_sum + b * C + c0 + e, t);
      }
    }
  }
}

// bf16 fast path.  The kernel above holds 8 channels x (9 taps + 9 window vectors) per thread = 177+ registers, one
// 192-thread block per SM, so HBM latency is exposed.  Here a lane owns ONE channel pair (a 32-bit load = both
// channels of a pixel, one packed fp32x2 FMA = both channels of a tap, lnx_dwconv_bf16.cu has the rationale) and
// R consecutive output rows: the window is (R-1)*STRIDE+3 input rows x 3 columns of already unpacked float2, the
// next column is loaded one iteration ahead as raw words, and each input row is read from L1/L2 (R+2)/R times
// instead of 3.  ~100 registers -> five 128-thread blocks per SM; a warp reads / writes 128 contiguous bytes per
// pixel.  Pool partials are folded over the block's row slots in shared memory, then one coalesced atomic per
// (block, channel).
template <int STRIDE, int R>
__global__ void __launch_bounds__(128, 5)
    dwconv3_x2_kernel(const bf16* __restrict__ x, const float* __restrict__ w9c, const float* __restrict__ scale, const float* __restrict__ shift,
                      bf16* __restrict__ y, float* __restrict__ pool_sum, int H, int W, int C, int pad_t, int pad_l, int Ho, int Wo, int act) {
  constexpr int NR = (R - 1) * STRIDE + 3;
  extern __shared__ float s_pool[];  // [blockDim.y][2 * blockDim.x] when pool_sum
  const int c0 = (blockIdx.z * blockDim.x + threadIdx.x) * 2;
  const long long b = blockIdx.y;
  const int ho0 = (blockIdx.x * blockDim.y + threadIdx.y) * R;
  float2 wr[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) wr[t] = *reinterpret_cast<const float2*>(w9c + t * C + c0);
  const float2 sc = scale ? *reinterpret_cast<const float2*>(scale + c0) : make_float2(1.f, 1.f);
  const float2 sh = shift ? *reinterpret_cast<const float2*>(shift + c0) : make_float2(0.f, 0.f);
  float2 psum = make_float2(0.f, 0.f);
  if (ho0 < Ho) {
    const int hbase = ho0 * STRIDE - pad_t;
    const uint32_t* xb = reinterpret_cast<const uint32_t*>(x + b * (long long)H * W * C + c0);
    const int rs = W * (C / 2), ps = C / 2;  // row / pixel strides in 32-bit words
    unsigned rowmask = 0;
#pragma unroll
    for (int r = 0; r < NR; ++r)
      if (hbase + r >= 0 && hbase + r < H) rowmask |= 1u << r;
    auto load_col = [&](int w, uint32_t* col) {
      const bool wok = w >= 0 && w < W;
      const uint32_t* p = xb + (long long)hbase * rs + (long long)w * ps;
#pragma unroll
      for (int r = 0; r < NR; ++r) col[r] = (wok && ((rowmask >> r) & 1u)) ? __ldg(p + (long long)r * rs) : 0u;
    };
    auto unpack = [](uint32_t v) { return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u)); };
    float2 win[3][NR];       // [column slot][row]; slot (wo + dw) % 3 holds input column wo * STRIDE - pad_l + dw (stride 1)
    uint32_t pre[STRIDE][NR];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      load_col(k - pad_l, pre[0]);
#pragma unroll
      for (int r = 0; r < NR; ++r) win[k][r] = unpack(pre[0][r]);
    }
    bf16* yb = y + ((b * Ho + ho0) * (long long)Wo) * C + c0;
    for (int wo3 = 0; wo3 < Wo; wo3 += 3) {
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int wo = wo3 + u;
        if (wo < Wo) {
          const int w0 = wo * STRIDE - pad_l;
          if (wo + 1 < Wo) {
#pragma unroll
            for (int k = 0; k < STRIDE; ++k) load_col(w0 + 3 + k, pre[k]);
          }
          // column slots of this output: stride 1 rotates one slot per output, stride 2 two slots per output
          const int s0 = (u * STRIDE) % 3, s1 = (u * STRIDE + 1) % 3, s2 = (u * STRIDE + 2) % 3;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (ho0 + r < Ho) {
              float2 acc = make_float2(0.f, 0.f);
#pragma unroll
              for (int dh = 0; dh < 3; ++dh) {
                acc = __ffma2_rn(win[s0][r * STRIDE + dh], wr[dh * 3 + 0], acc);
                acc = __ffma2_rn(win[s1][r * STRIDE + dh], wr[dh * 3 + 1], acc);
                acc = __ffma2_rn(win[s2][r * STRIDE + dh], wr[dh * 3 + 2], acc);
              }
              acc = __ffma2_rn(acc, sc, sh);
              if (act == 1) {  // x * sigmoid(x) with ex2 / rcp (the tanh.approx form costs ~1/8 bf16 ulp more; the kernel is HBM bound)
                acc.x = __fdividef(acc.x, 1.f + __expf(-acc.x));
                acc.y = __fdividef(acc.y, 1.f + __expf(-acc.y));
              }
              const __nv_bfloat162 o = __floats2bfloat162_rn(acc.x, acc.y);
              const uint32_t ow = *reinterpret_cast<const uint32_t*>(&o);
              *reinterpret_cast<uint32_t*>(yb + ((long long)r * Wo + wo) * C) = ow;
              const float2 of = unpack(ow);  // pool what the next layer will read (the rounded value)
              psum.x += of.x;
              psum.y += of.y;
            }
          }
          // retire the oldest column(s): the slots the next output no longer needs take the prefetched columns
#pragma unroll
          for (int k = 0; k < STRIDE; ++k) {
            const int slot = (u * STRIDE + k) % 3;
#pragma unroll
            for (int r = 0; r < NR; ++r) win[slot][r] = unpack(pre[k][r]);
          }
        }
      }
    }
  }
  if (pool_sum) {
    const int tx2 = 2 * blockDim.x;
    s_pool[threadIdx.y * tx2 + 2 * threadIdx.x] = psum.x;
    s_pool[threadIdx.y * tx2 + 2 * threadIdx.x + 1] = psum.y;
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    for (int i = tid; i < tx2; i += blockDim.x * blockDim.y) {
      float t = 0.f;
      for (int r = 0; r < (int)blockDim.y; ++r) t += s_pool[r * tx2 + i];
      atomicAdd(pool_sum + b * C + blockIdx.z * tx2 + i, t);
    }
  }
}

// ---------------------------------------------------------------- x * sigmoid(gate[b, c])
// grid (pixel chunks, B); block (C / V, rows): a thread owns one channel vector, evaluates its V sigmoids once and streams
// the chunk's pixels through them (one 16-byte load, V multiplies, one 16-byte store per pixel).
template <typename T>
__global__ void se_scale_kernel(const T* __restrict__ x, const float* __restrict__ gate, T* __restrict__ y, int HW, int C, int chunk) {
  constexpr int V = Vec16<T>::N;
  const long long b = blockIdx.y;
  const int c0 = threadIdx.x * V;
  float g[V];
#pragma unroll
  for (int e = 0; e < V; ++e) g[e] = 1.f / (1.f + expf(-gate[b * C + c0 + e]));
  const int p0 = blockIdx.x * chunk, p1 = min(HW, p0 + chunk);
  const long long base = b * HW * (long long)C + c0;
#pragma unroll 4
  for (int p = p0 + threadIdx.y; p < p1; p += blockDim.y) {
    Vec16<T> v = ld16(x + base + (long long)p * C);
#pragma unroll
    for (int e = 0; e < V; ++e) v.set(e, v.get(e) * g[e]);
    st16(y + base + (long long)p * C, v);
  }
}

// ---------------------------------------------------------------- attention with additive bias, generic head dim
constexpr int AKT = 32;  // keys per shared tile
template <typename T, int HDIM>
__global__ void __launch_bounds__(64) attn_bias_fwd_kernel(const T* __restrict__ qkv, const float* __restrict__ bias, T* __restrict__ out,
                                                           int heads, int N, float scale) {
  __shared__ __align__(16) float Ks[AKT * HDIM];
  __shared__ __align__(16) float Vs[AKT * HDIM];
  const int bh = blockIdx.y;
  const int b = bh / heads, h = bh % heads;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < N;
  const long long row_stride = 3LL * heads * HDIM;  // elements per token in qkv
  const T* base = qkv + (long long)b * N * row_stride + (long long)h * HDIM;
  float qr[HDIM], o[HDIM];
#pragma unroll
  for (int d = 0; d < HDIM; ++d) {
    qr[d] = valid ? to_f32(base[(long long)i * row_stride + d]) * scale : 0.f;
    o[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  const float* brow = bias ? bias + ((long long)h * N + (valid ? i : 0)) * N : nullptr;
  for (int j0 = 0; j0 < N; j0 += AKT) {
    __syncthreads();
    for (int t = threadIdx.x; t < AKT * HDIM; t += blockDim.x) {
      const int r = t / HDIM, d = t % HDIM;
      float kv = 0.f, vv = 0.f;
      if (j0 + r < N) {
        const T* tok = base + (long long)(j0 + r) * row_stride;
        kv = to_f32(tok[(long long)heads * HDIM + d]);
        vv = to_f32(tok[2LL * heads * HDIM + d]);
      }
      Ks[t] = kv;
      Vs[t] = vv;
    }
    __syncthreads();
    const int jn = min(AKT, N - j0);
    for (int jc = 0; jc < jn; ++jc) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < HDIM; ++d) a = fmaf(qr[d], Ks[jc * HDIM + d], a);
      if (brow) a += brow[j0 + jc];
      const float mn = fmaxf(m, a);
      const float corr = __expf(m - mn);
      const float pe = __expf(a - mn);
      l = l * corr + pe;
#pragma unroll
      for (int d = 0; d < HDIM; ++d) o[d] = fmaf(o[d], corr, pe * Vs[jc * HDIM + d]);
      m = mn;
    }
  }
  if (valid) {
    const float inv = 1.f / l;
    T* dst = out + ((long long)b * N + i) * heads * HDIM + (long long)h * HDIM;
#pragma unroll
    for (int d = 0; d < HDIM; ++d) dst[d] = from_f32<T>(o[d] * inv);
  }
}

// short sequences (N <= 64: the last stage, 49 patches + extras) at any head_dim that is a multiple of 32: one CTA per
// (batch, head) keeps K and V in shared memory as fp32 (K rows padded by 4 floats: conflict-free 16-byte reads with a lane
// per key), a warp per query row: lanes = keys for q k^T + bias and the softmax (two keys per lane), lanes = head dims for
// P V.  The generic kernel above holds a whole q and o row per thread (2 x head_dim registers) and runs at one warp per SM
// sub-partition for these shapes.
template <typename T>
__global__ void __launch_bounds__(128) attn_small_kernel(const T* __restrict__ qkv, const float* __restrict__ bias, T* __restrict__ out, int heads,
                                                         int N, int hd, float scale) {
  extern __shared__ __align__(16) float sm_attn[];
  constexpr int V = Vec16<T>::N;
  const int ks = hd + 4;
  float* Ks = sm_attn;
  float* Vs = Ks + N * ks;
  float* sq = Vs + N * hd;
  float* sp = sq + 16 * hd;  // [4 warps][4 rows][hd] q rows, then [4][4][64] probabilities
  const int bh = blockIdx.x;
  const int b = bh / heads, h = bh - b * heads;
  const long long row_stride = 3LL * heads * hd;
  const T* base = qkv + (long long)b * N * row_stride + (long long)h * hd;
  const int vec_per_row = hd / V;
  for (int t = threadIdx.x; t < N * vec_per_row; t += blockDim.x) {
    const int j = t / vec_per_row, dv = (t - j * vec_per_row) * V;
    const T* tok = base + (long long)j * row_stride + dv;
    const Vec16<T> kv = ld16(tok + (long long)heads * hd), vv = ld16(tok + 2LL * heads * hd);
#pragma unroll
    for (int e = 0; e < V; ++e) {
      Ks[j * ks + dv + e] = kv.get(e);
      Vs[j * hd + dv + e] = vv.get(e);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int R = 4;  // query rows per warp pass: every K / V value read from shared memory feeds R rows
  float* q = sq + warp * R * hd;
  float* pw = sp + warp * R * 64;
  const int j0 = lane, j1 = lane + 32;
  const float* k0 = Ks + min(j0, N - 1) * ks;
  const float* k1 = Ks + min(j1, N - 1) * ks;
  // the q rows and bias values of the NEXT pass are fetched while this one is computed (global latency per pass otherwise)
  float qn[R][4], bn0[R], bn1[R];
  auto fetch = [&](int i0) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = i0 + r;
#pragma unroll
      for (int t = 0; t < 4; ++t) qn[r][t] = (i < N && lane + 32 * t < hd) ? to_f32(base[(long long)i * row_stride + lane + 32 * t]) * scale : 0.f;
      bn0[r] = bn1[r] = 0.f;
      if (bias && i < N) {
        const float* brow = bias + ((long long)h * N + i) * N;
        if (j0 < N) bn0[r] = brow[j0];
        if (j1 < N) bn1[r] = brow[j1];
      }
    }
  };
  fetch(warp * R);
  for (int i0 = warp * R; i0 < N; i0 += 4 * R) {
    float bc0[R], bc1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (lane + 32 * t < hd) q[r * hd + lane + 32 * t] = qn[r][t];
      bc0[r] = bn0[r];
      bc1[r] = bn1[r];
    }
    __syncwarp();
    fetch(i0 + 4 * R);
    float s0[R], s1[R];
#pragma unroll
    for (int r = 0; r < R; ++r) s0[r] = s1[r] = 0.f;
#pragma unroll 2
    for (int d = 0; d < hd; d += 4) {
      const float4 a = *reinterpret_cast<const float4*>(k0 + d);
      const float4 c = *reinterpret_cast<const float4*>(k1 + d);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float4 qv = *reinterpret_cast<const float4*>(q + r * hd + d);
        s0[r] = fmaf(qv.x, a.x, s0[r]); s0[r] = fmaf(qv.y, a.y, s0[r]); s0[r] = fmaf(qv.z, a.z, s0[r]); s0[r] = fmaf(qv.w, a.w, s0[r]);
        s1[r] = fmaf(qv.x, c.x, s1[r]); s1[r] = fmaf(qv.y, c.y, s1[r]); s1[r] = fmaf(qv.z, c.z, s1[r]); s1[r] = fmaf(qv.w, c.w, s1[r]);
      }
    }
    float inv[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float x0 = j0 < N ? s0[r] + bc0[r] : -INFINITY, x1 = j1 < N ? s1[r] + bc1[r] : -INFINITY;
      const float m = warp_max(fmaxf(x0, x1));
      const float e0 = j0 < N ? (sizeof(T) == 4 ? expf(x0 - m) : __expf(x0 - m)) : 0.f;
      const float e1 = j1 < N ? (sizeof(T) == 4 ? expf(x1 - m) : __expf(x1 - m)) : 0.f;
      inv[r] = 1.f / warp_sum(e0 + e1);
      pw[r * 64 + j0] = e0;
      pw[r * 64 + j1] = e1;
    }
    __syncwarp();
    float o[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int t = 0; t < 4; ++t) o[r][t] = 0.f;
    for (int j = 0; j < N; ++j) {
      float vv[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) vv[t] = lane + 32 * t < hd ? Vs[j * hd + lane + 32 * t] : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float pj = pw[r * 64 + j];
#pragma unroll
        for (int t = 0; t < 4; ++t) o[r][t] = fmaf(pj, vv[t], o[r][t]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = i0 + r;
      if (i < N) {
        T* dst = out + ((long long)b * N + i) * heads * hd + (long long)h * hd;
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (lane + 32 * t < hd) dst[lane + 32 * t] = from_f32<T>(o[r][t] * inv[r]);
      }
    }
    __syncwarp();
  }
}

template <typename T>
int launch_attn_small(const void* qkv, const float* bias, void* out, int B, int heads, int N, int hd, float scale, cudaStream_t st) {
  const size_t smem = ((size_t)N * (hd + 4) + (size_t)N * hd + 16 * hd + 16 * 64) * sizeof(float);
  static size_t smem_set = 48 * 1024;
  if (smem > smem_set) {
    if (cudaFuncSetAttribute(attn_small_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
      cudaGetLastError();
      return LNX_ERR_UNSUPPORTED;
    }
    smem_set = smem;
  }
  attn_small_kernel<T><<<B * heads, 128, smem, st>>>((const T*)qkv, bias, (T*)out, heads, N, hd, scale);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

inline int grid_for(long long total, int threads) { return (int)max(1LL, min((long long)kNumSMs * 16, (total + threads - 1) / threads)); }

}  // namespace

extern "C" int lnx_im2col3x3(const void* x, int x_is_nchw_f32, void* out, int B, int H, int W, int C, int stride, int Ho, int Wo, int Kpad,
                             int dtype, lnx_stream_t s) {
  LNX_REQUIRE(x && out, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && (stride == 1 || stride == 2) && Kpad >= 9 * C, LNX_ERR_SHAPE);
  LNX_REQUIRE(Ho == (H + 2 - 3) / stride + 1 && Wo == (W + 2 - 3) / stride + 1, LNX_ERR_SHAPE);
  cudaStream_t st = (cudaStream_t)s;
  if (x_is_nchw_f32) {
    const long long total = (long long)B * Ho * Wo;
    if (dtype == LNX_F32) im2col3_nchw_kernel<float><<<grid_for(total, 128), 128, 0, st>>>((const float*)x, (float*)out, B, H, W, C, stride, Ho, Wo, Kpad);
    else if (dtype == LNX_BF16 && C == 3 && Kpad == 32 && lnx_aligned16(out))
      im2col3_rgb_bf16_kernel<<<grid_for(total, 256), 256, 0, st>>>((const float*)x, (bf16*)out, B, H, W, stride, Ho, Wo);
    else if (dtype == LNX_BF16) im2col3_nchw_kernel<bf16><<<grid_for(total, 128), 128, 0, st>>>((const float*)x, (bf16*)out, B, H, W, C, stride, Ho, Wo, Kpad);
    else return LNX_ERR_DTYPE;
  } else {
    const int V = dtype == LNX_F32 ? 4 : 8;
    LNX_REQUIRE(dtype == LNX_F32 || dtype == LNX_BF16, LNX_ERR_DTYPE);
    if (C % V != 0 || Kpad != 9 * C) {
      const long long total = (long long)B * Ho * Wo * Kpad;
      if (dtype == LNX_F32) im2col3_nhwc_scalar_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)x, (float*)out, B, H, W, C, stride, Ho, Wo, Kpad);
      else im2col3_nhwc_scalar_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)x, (bf16*)out, B, H, W, C, stride, Ho, Wo, Kpad);
    } else {
      LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(out), LNX_ERR_ALIGN);
      const long long total = (long long)B * Ho * Wo * 9 * (C / V);
      if (dtype == LNX_F32) im2col3_nhwc_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)x, (float*)out, B, H, W, C, stride, Ho, Wo, Kpad);
      else im2col3_nhwc_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)x, (bf16*)out, B, H, W, C, stride, Ho, Wo, Kpad);
    }
  }
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_maxpool3s2(const void* x, void* y, int B, int H, int W, int C, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(x && y, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(y), LNX_ERR_ALIGN);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_F32) {
    LNX_REQUIRE(C % 4 == 0, LNX_ERR_SHAPE);
    maxpool3s2_kernel<float><<<grid_for((long long)B * Ho * Wo * (C / 4), 256), 256, 0, st>>>((const float*)x, (float*)y, B, H, W, C, Ho, Wo);
  } else if (dtype == LNX_BF16) {
    LNX_REQUIRE(C % 8 == 0, LNX_ERR_SHAPE);
    maxpool3s2_kernel<bf16><<<grid_for((long long)B * Ho * Wo * (C / 8), 256), 256, 0, st>>>((const bf16*)x, (bf16*)y, B, H, W, C, Ho, Wo);
  } else
    return LNX_ERR_DTYPE;
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_dwconv3_fwd(const void* x, const float* w9c, const float* scale, const float* shift, void* y, float* pool_sum, int B, int H,
                               int W, int C, int stride, int pad_t, int pad_l, int Ho, int Wo, int act, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(x && w9c && y, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && Ho > 0 && Wo > 0 && (stride == 1 || stride == 2), LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(y), LNX_ERR_ALIGN);
  const int V = dtype == LNX_F32 ? 4 : 8;
  LNX_REQUIRE(dtype == LNX_F32 || dtype == LNX_BF16, LNX_ERR_DTYPE);
  LNX_REQUIRE(C % V == 0 && C / V <= 1024, LNX_ERR_SHAPE);
  cudaStream_t st = (cudaStream_t)s;
  if (dtype == LNX_BF16) {  // channel-pair kernel (FFMA2)
    const int cp = C / 2;
    int tx = 0;
    for (int d = 128; d >= 1 && !tx; --d)
      if (cp % d == 0 && (d % 32 == 0 || d == cp)) tx = d;
    if (!tx)
      for (int d = 128; d >= 1 && !tx; --d)
        if (cp % d == 0) tx = d;
    const int R = stride == 1 ? 4 : 2;
    const int slots = (Ho + R - 1) / R;
    int ty = max(1, min(128 / tx, slots));
    while (ty > 1 && ((slots + ty - 1) / ty) * ty - slots > slots / 8) --ty;
    dim3 grid((slots + ty - 1) / ty, B, cp / tx), block(tx, ty);
    const size_t smem = pool_sum ? (size_t)ty * 2 * tx * sizeof(float) : 0;
    if (stride == 1)
      dwconv3_x2_kernel<1, 4><<<grid, block, smem, st>>>((const bf16*)x, w9c, scale, shift, (bf16*)y, pool_sum, H, W, C, pad_t, pad_l, Ho, Wo, act);
    else
      dwconv3_x2_kernel<2, 2><<<grid, block, smem, st>>>((const bf16*)x, w9c, scale, shift, (bf16*)y, pool_sum, H, W, C, pad_t, pad_l, Ho, Wo, act);
    LNX_CHECK_LAUNCH();
    return LNX_OK;
  }
  const int tx = C / V;
  const int ty = max(1, min(8, 256 / tx));
  dim3 grid((Ho + ty - 1) / ty, B), block(tx, ty);
  const size_t smem = pool_sum ? (size_t)ty * C * sizeof(float) : 0;
#define LNX_DW3(T, S) dwconv3_kernel<T, S><<<grid, block, smem, st>>>((const T*)x, w9c, scale, shift, (T*)y, pool_sum, H, W, C, pad_t, pad_l, Ho, Wo, act)
  if (stride == 1) LNX_DW3(float, 1); else LNX_DW3(float, 2);
#undef LNX_DW3
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

extern "C" int lnx_se_scale(const void* x, const float* gate, void* y, int B, int HW, int C, int dtype, lnx_stream_t s) {
  LNX_REQUIRE(x && gate && y, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && HW > 0 && C > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(lnx_aligned16(x) && lnx_aligned16(y), LNX_ERR_ALIGN);
  cudaStream_t st = (cudaStream_t)s;
  LNX_REQUIRE(dtype == LNX_F32 || dtype == LNX_BF16, LNX_ERR_DTYPE);
  const int V = dtype == LNX_F32 ? 4 : 8;
  LNX_REQUIRE(C % V == 0 && C / V <= 1024 && B <= 65535, LNX_ERR_SHAPE);
  const int tx = C / V, ty = max(1, 256 / tx);
  // ~8 chunks per SM over the whole batch, at least 4 pixels per thread
  int chunks = max(1, min((HW + 4 * ty - 1) / (4 * ty), (8 * kNumSMs + B - 1) / B));
  const int chunk = (HW + chunks - 1) / chunks;
  chunks = (HW + chunk - 1) / chunk;
  dim3 grid(chunks, B), block(tx, ty);
  if (dtype == LNX_F32) se_scale_kernel<float><<<grid, block, 0, st>>>((const float*)x, gate, (float*)y, HW, C, chunk);
  else se_scale_kernel<bf16><<<grid, block, 0, st>>>((const bf16*)x, gate, (bf16*)y, HW, C, chunk);
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}

int lnx_attn_bias_fwd_tc2(const void* qkv, const float* bias, void* out, float* lse, int B, int heads, int N, int hd, float scale, cudaStream_t st);

extern "C" int lnx_attn_bias_fwd(const void* qkv, const float* bias, void* out, int B, int heads, int N, int hd, float scale, int dtype,
                                 lnx_stream_t s) {
  LNX_REQUIRE(qkv && out, LNX_ERR_NULL);
  LNX_REQUIRE(B > 0 && heads > 0 && N > 0, LNX_ERR_SHAPE);
  LNX_REQUIRE(dtype == LNX_F32 || dtype == LNX_BF16, LNX_ERR_DTYPE);
  if (dtype == LNX_BF16) {  // tcgen05 path: head_dim <= 64, N <= 240 (persistent pipelined kernel of lnx_attn_tc2.cu)
    const int r = lnx_attn_bias_fwd_tc2(qkv, bias, out, nullptr, B, heads, N, hd, scale, (cudaStream_t)s);
    if (r != LNX_ERR_UNSUPPORTED) return r;
    if (getenv("LNX_ATTN_NO_FALLBACK") && hd <= 64 && N <= 240) return LNX_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)s;
  if (N <= 64 && hd % 32 == 0 && hd <= 128 && lnx_aligned16(qkv)) {  // K / V resident in shared memory, a warp per query row
    const int r = dtype == LNX_F32 ? launch_attn_small<float>(qkv, bias, out, B, heads, N, hd, scale, st)
                                   : launch_attn_small<bf16>(qkv, bias, out, B, heads, N, hd, scale, st);
    if (r != LNX_ERR_UNSUPPORTED) return r;
  }
  dim3 grid((N + 63) / 64, B * heads);
#define LNX_AB(HDIM)                                                                                                                   \
  case HDIM:                                                                                                                           \
    if (dtype == LNX_F32) attn_bias_fwd_kernel<float, HDIM><<<grid, 64, 0, st>>>((const float*)qkv, bias, (float*)out, heads, N, scale); \
    else attn_bias_fwd_kernel<bf16, HDIM><<<grid, 64, 0, st>>>((const bf16*)qkv, bias, (bf16*)out, heads, N, scale);                     \
    break;
  switch (hd) {
    LNX_AB(16) LNX_AB(32) LNX_AB(48) LNX_AB(64) LNX_AB(96) LNX_AB(128)
    default: return LNX_ERR_UNSUPPORTED;
  }
#undef LNX_AB
  LNX_CHECK_LAUNCH();
  return LNX_OK;
}
