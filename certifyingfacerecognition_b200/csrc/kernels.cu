// HBM-bound and small kernels of the certification path.  All are coalesced / 128-bit vectorised over the
// NHWC channel dimension; reductions go registers -> shared -> one global atomic per (block, channel).
#include "kernels.cuh"
#include "ptx.cuh"

#include <cstdio>
#include <type_traits>
#include <cstdlib>

namespace cfr {
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define CFR_LAUNCH_CHECK(name)                                                   \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      set_error("%s launch failed: %s", name, cudaGetErrorString(e__));          \
      return 4;                                                                  \
    }                                                                            \
    count_launch();                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based RNG (Salmon et al. 2011): counter = (sample index, draw), key = seed.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = (static_cast<float>(a) + 1.0f) * 2.3283064365386963e-10f;   // (0,1]
  const float u2 = static_cast<float>(b) * 2.3283064365386963e-10f;            // [0,1)
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincosf(6.283185307179586f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

__global__ void k_noise_project(const float* __restrict__ z, const float* __restrict__ x,
                                const float* __restrict__ sigma, int sigma_len, const float* __restrict__ noise_in,
                                const float* __restrict__ dir_mat, const float* __restrict__ w_avg, float psi,
                                unsigned long long seed, unsigned long long sample_offset, int b,
                                float* __restrict__ noise_out, float* __restrict__ wp2) {
  const int s = blockIdx.x;
  __shared__ float p[5];
  if (threadIdx.x == 0) {
    float nz[8];
    if (noise_in != nullptr) {
      for (int k = 0; k < 5; ++k) nz[k] = noise_in[s * 5 + k];
    } else {
      const unsigned long long idx = sample_offset + s;
      const uint2 key = make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
      for (int d = 0; d < 2; ++d) {
        const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), d, 0), key);
        const float2 g0 = box_muller(r.x, r.y), g1 = box_muller(r.z, r.w);
        nz[4 * d] = g0.x; nz[4 * d + 1] = g0.y; nz[4 * d + 2] = g1.x; nz[4 * d + 3] = g1.y;
      }
      for (int k = 0; k < 5; ++k) nz[k] *= sigma[sigma_len == 1 ? 0 : k];
    }
    for (int k = 0; k < 5; ++k) {
      if (noise_out != nullptr) noise_out[s * 5 + k] = nz[k];
      p[k] = x[k] + nz[k];
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < 512; d += blockDim.x) {
    float pert = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) pert += p[k] * dir_mat[k * 512 + d];
    const float w = z[d] + pert;
    const float wa = w_avg[d];
    wp2[(static_cast<size_t>(s) * 2 + 0) * 512 + d] = wa + (w - wa) * psi;
    wp2[(static_cast<size_t>(s) * 2 + 1) * 512 + d] = wa + (w - wa) * 1.0f;
  }
}

int launch_noise_project(const float* z, const float* x, const float* sigma, int sigma_len, const float* noise_in,
                         const float* dir_mat, const float* w_avg, float psi, unsigned long long seed,
                         unsigned long long sample_offset, int b, float* noise_out, float* wp2, cudaStream_t st) {
  if (b <= 0) return 0;
  k_noise_project<<<b, 128, 0, st>>>(z, x, sigma, sigma_len, noise_in, dir_mat, w_avg, psi, seed, sample_offset, b,
                                     noise_out, wp2);
  CFR_LAUNCH_CHECK("noise_project");
  return 0;
}

__global__ void k_truncate(const float* __restrict__ w, const float* __restrict__ w_avg, float psi, int b,
                           float* __restrict__ wp2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b * 512) return;
  const int s = i / 512, d = i % 512;
  const float wa = w_avg[d], v = w[i];
  wp2[(static_cast<size_t>(s) * 2 + 0) * 512 + d] = wa + (v - wa) * psi;
  wp2[(static_cast<size_t>(s) * 2 + 1) * 512 + d] = wa + (v - wa) * 1.0f;
}
int launch_truncate(const float* w, const float* w_avg, float psi, int b, float* wp2, cudaStream_t st) {
  if (b <= 0) return 0;
  k_truncate<<<(b * 512 + 255) / 256, 256, 0, st>>>(w, w_avg, psi, b, wp2);
  CFR_LAUNCH_CHECK("truncate");
  return 0;
}

// ------------------------------------------------------------------------------------------
// styles[b][row] = <wp2[b][variant(row)], W[row]> / sqrt(512) + bias[row]
// A block owns 32 consecutive style rows and a slab of 32 samples: both 32 x 512 operands are staged in shared memory
// (k-chunks of 64), thread (sample, row group) accumulates 4 rows in registers -- no shuffles, every weight element is
// read once per 32 samples.  (The first version used one warp per row with a shuffle reduction per sample: 1.6 us/sample.)
// ------------------------------------------------------------------------------------------
constexpr int kStRows = 64, kStSamples = 64, kStK = 32;
__global__ void __launch_bounds__(256) k_styles(const float* __restrict__ wp2, const float* __restrict__ w_style,
                                                const float* __restrict__ b_style, int rows, int rows_trunc, int b,
                                                float* __restrict__ styles) {
  // 64 rows x 64 samples per block, 4 x 4 per thread (8 shared loads per 16 FMAs; the 4 x 1 version was LDS-bound)
  __shared__ float ws[kStRows][kStK + 1];
  __shared__ float xs[kStSamples][kStK + 1];
  const int row0 = blockIdx.x * kStRows, s0 = blockIdx.y * kStSamples;
  // rows_trunc is a multiple of 64 (launch_styles checks), so a block never mixes the two w variants
  const int variant = row0 < rows_trunc ? 0 : 1;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;        // samples tx + 16 i, rows ty * 4 + j
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < 512; k0 += kStK) {
    for (int i = threadIdx.x; i < kStRows * kStK; i += 256) {
      const int r = i / kStK, k = i % kStK;
      ws[r][k] = row0 + r < rows ? __ldg(w_style + static_cast<size_t>(row0 + r) * 512 + k0 + k) : 0.f;
      xs[r][k] = s0 + r < b ? __ldg(wp2 + (static_cast<size_t>(s0 + r) * 2 + variant) * 512 + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kStK; ++k) {
      float x[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = xs[tx + 16 * i][k];
#pragma unroll
      for (int j = 0; j < 4; ++j) w[j] = ws[ty * 4 + j][k];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int smp = s0 + tx + 16 * i;
    if (smp >= b) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = row0 + ty * 4 + j;
      if (row < rows) styles[static_cast<size_t>(smp) * rows + row] = acc[i][j] * 0.044194173824159216f + b_style[row];
    }
  }
}
int launch_styles(const float* wp2, const float* w_style, const float* b_style, int rows, int rows_trunc, int b,
                  float* styles, cudaStream_t st) {
  if (rows_trunc % kStRows != 0) { set_error("styles: rows_trunc=%d must be a multiple of %d", rows_trunc, kStRows); return 2; }
  dim3 grid((rows + kStRows - 1) / kStRows, (b + kStSamples - 1) / kStSamples);
  k_styles<<<grid, 256, 0, st>>>(wp2, w_style, b_style, rows, rows_trunc, b, styles);
  CFR_LAUNCH_CHECK("styles");
  return 0;
}

// ------------------------------------------------------------------------------------------
__global__ void k_layer0(const float* __restrict__ xhat0, const float* __restrict__ styles, int style_stride,
                         int style_off, int b, __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over b*16*512
  if (i >= b * 16 * 512) return;
  const int c = i % 512, pix = (i / 512) % 16, s = i / (512 * 16);
  const float s0 = styles[static_cast<size_t>(s) * style_stride + style_off + c];
  const float s1 = styles[static_cast<size_t>(s) * style_stride + style_off + 512 + c];
  out[i] = __float2half_rn(xhat0[pix * 512 + c] * (s0 + 1.f) + s1);
}
int launch_layer0(const float* xhat0, const float* styles, int style_stride, int style_off, int b, __half* out,
                  cudaStream_t st) {
  k_layer0<<<(b * 16 * 512 + 255) / 256, 256, 0, st>>>(xhat0, styles, style_stride, style_off, b, out);
  CFR_LAUNCH_CHECK("layer0");
  return 0;
}
// Split-precision variant: out [b,16,3*512] = [hi | lo | hi] with hi = fp16(x), lo = fp16(x - hi), the operand layout of
// the hi/lo-split convolutions (cfr_conv_desc.kSplit == 3).
__global__ void k_layer0_split(const float* __restrict__ xhat0, const float* __restrict__ styles, int style_stride,
                               int style_off, int b, __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over b*16*512
  if (i >= b * 16 * 512) return;
  const int c = i % 512, pix = i / 512, s = i / (512 * 16);
  const float s0 = styles[static_cast<size_t>(s) * style_stride + style_off + c];
  const float s1 = styles[static_cast<size_t>(s) * style_stride + style_off + 512 + c];
  const float v = xhat0[(pix % 16) * 512 + c] * (s0 + 1.f) + s1;
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  __half* o = out + static_cast<size_t>(pix) * 1536 + c;
  o[0] = hi;
  o[512] = lo;
  o[1024] = hi;
}
int launch_layer0_split(const float* xhat0, const float* styles, int style_stride, int style_off, int b, __half* out,
                        cudaStream_t st) {
  k_layer0_split<<<(b * 16 * 512 + 255) / 256, 256, 0, st>>>(xhat0, styles, style_stride, style_off, b, out);
  CFR_LAUNCH_CHECK("layer0_split");
  return 0;
}

// ------------------------------------------------------------------------------------------
// Mapping network Z -> W (generate_data.py path).  8 dense 512x512 layers in fp32 on the CUDA cores: 4.2 MFLOP per
// latent, weights (8 MB) L2-resident and read coalesced through the transposed layout; one block = kMapRows latents,
// thread o owns output feature o of every layer.
// ------------------------------------------------------------------------------------------
constexpr int kMapRows = 8;
__global__ void __launch_bounds__(512) k_mapping(const float* __restrict__ z, const float* __restrict__ wt,
                                                 const float* __restrict__ bias, int b, float* __restrict__ w_out) {
  __shared__ float xs[2][kMapRows][512];
  __shared__ float red[kMapRows][16];
  const int o = threadIdx.x, r0 = blockIdx.x * kMapRows;
  const int lane = o & 31, warp = o >> 5;
  // PixelNormLayer: x / sqrt(mean(x^2) + 1e-8)
  float v[kMapRows];
#pragma unroll
  for (int s = 0; s < kMapRows; ++s) {
    v[s] = (r0 + s < b) ? z[static_cast<size_t>(r0 + s) * 512 + o] : 0.f;
    float q = v[s] * v[s];
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    if (lane == 0) red[s][warp] = q;
  }
  __syncthreads();
#pragma unroll
  for (int s = 0; s < kMapRows; ++s) {
    float q = 0.f;
    for (int w = 0; w < 16; ++w) q += red[s][w];
    xs[0][s][o] = v[s] / sqrtf(q * (1.0f / 512.0f) + 1e-8f);
  }
  __syncthreads();
  int cur = 0;
  for (int l = 0; l < 8; ++l) {
    const float* wl = wt + static_cast<size_t>(l) * 512 * 512 + o;
    float acc[kMapRows];
#pragma unroll
    for (int s = 0; s < kMapRows; ++s) acc[s] = 0.f;
#pragma unroll 4
    for (int i = 0; i < 512; ++i) {
      const float wv = __ldg(wl + static_cast<size_t>(i) * 512);
#pragma unroll
      for (int s = 0; s < kMapRows; ++s) acc[s] = fmaf(xs[cur][s][i], wv, acc[s]);
    }
    const float bv = bias[l * 512 + o];
#pragma unroll
    for (int s = 0; s < kMapRows; ++s) {
      const float t = acc[s] + bv;
      xs[cur ^ 1][s][o] = t >= 0.f ? t : 0.2f * t;
    }
    __syncthreads();
    cur ^= 1;
  }
#pragma unroll
  for (int s = 0; s < kMapRows; ++s)
    if (r0 + s < b) w_out[static_cast<size_t>(r0 + s) * 512 + o] = xs[cur][s][o];
}
int launch_mapping(const float* z, const float* wt, const float* bias, int b, float* w_out, cudaStream_t st) {
  if (b <= 0) return 0;
  k_mapping<<<(b + kMapRows - 1) / kMapRows, 512, 0, st>>>(z, wt, bias, b, w_out);
  CFR_LAUNCH_CHECK("mapping");
  return 0;
}

// ------------------------------------------------------------------------------------------
// blur [1,2,1]x[1,2,1]/16 (zero pad) + noise*w + bias + lrelu(0.2) + per-(n,c) sum / sumsq
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void load8(const __half* p, float (&f)[8]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const __half2* h2 = reinterpret_cast<const __half2*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h2[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(__half* p, const float (&f)[8]) {
  uint4 o;
  __half2* h2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) h2[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = o;
}

__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// Block-level reduction of per-thread channel sums for the InstanceNorm statistics.  256 threads; thread t owns the
// 8 channels starting at (t % c8) * 8.  Warp shuffles, then one float slot per (warp-slice, channel) in `part`
// ([2][2048] floats), then a fixed-order sum -> Q43.20 -> one global 64-bit RED per channel: deterministic, and no
// 64-bit shared-memory atomics (those compile to ATOMS.CAST.SPIN loops; 64 contending threads made them the
// dominant cost of the blur pass).  Must be called by all 256 threads; `part` must not be in use.
__device__ __forceinline__ void block_flush_stats(float (&acc)[8], float (&acc2)[8], int c, int n,
                                                  stat_t* __restrict__ gsum, stat_t* __restrict__ gsq, float* part) {
  const int c8 = c >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int off = 16; off >= c8; off >>= 1) {             // lanes with equal lane % c8 own the same channels
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
      acc2[i] += __shfl_xor_sync(0xffffffffu, acc2[i], off);
    }
  }
  const int wpc = c8 > 32 ? c8 >> 5 : 1;                 // warps that together cover one channel set
  const int slices = 8 / wpc, slice = warp / wpc;
  const int ch = (threadIdx.x % c8) * 8;
  if (c8 >= 32 || lane < c8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      part[slice * c + ch + i] = acc[i];
      part[2048 + slice * c + ch + i] = acc2[i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < c; i += 256) {
    float a = 0.f, b = 0.f;
    for (int sl = 0; sl < slices; ++sl) {
      a += part[sl * c + i];
      b += part[2048 + sl * c + i];
    }
    atomicAdd(&gsum[n * c + i], stat_fx(a));
    atomicAdd(&gsq[n * c + i], stat_fx(b));
  }
}

template <int MODE, typename T = __half>
__global__ void __launch_bounds__(256) k_blur_act_stats(const T* __restrict__ raw, T* __restrict__ y, int h,
                                                        int w, int c, const float* __restrict__ noise,
                                                        const float* __restrict__ noise_w,
                                                        const float* __restrict__ bias, stat_t* __restrict__ gsum,
                                                        stat_t* __restrict__ gsq) {
  __shared__ float s_part[2 * 2048];
  const int n = blockIdx.y;
  const int c8 = c >> 3;
  const int ppb = 256 / c8;                 // pixels per block step (c8 <= 64)
  const int cg = threadIdx.x % c8, pl = threadIdx.x / c8;
  const int ch = cg * 8;
  float nw[8], bs[8], acc[8], acc2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    nw[i] = MODE == 0 ? noise_w[ch + i] : 0.f;
    bs[i] = MODE == 0 ? bias[ch + i] : 0.f;
    acc[i] = 0.f;
    acc2[i] = 0.f;
  }
  const int hw = h * w;
  const T* img = raw + static_cast<size_t>(n) * hw * c;
  if (pl < ppb) {
    for (int pix = blockIdx.x * ppb + pl; pix < hw; pix += gridDim.x * ppb) {
      float v[8];
      if (MODE == 0) {
        const int py = pix / w, px = pix - py * w;
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy) {
          const int yy = py + dy;
          if (yy < 0 || yy >= h) continue;
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int xx = px + dx;
            if (xx < 0 || xx >= w) continue;
            const float k = (dy == 0 ? 2.f : 1.f) * (dx == 0 ? 2.f : 1.f) * 0.0625f;
            float t[8];
            load8(img + (static_cast<size_t>(yy) * w + xx) * c + ch, t);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += k * t[i];
          }
        }
        const float nz = __ldg(&noise[pix]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = v[i] + nz * nw[i] + bs[i];
          v[i] = t >= 0.f ? t : 0.2f * t;
        }
        store8(y + (static_cast<size_t>(n) * hw + pix) * c + ch, v);
      } else {
        load8(img + static_cast<size_t>(pix) * c + ch, v);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i] += v[i];
        acc2[i] += v[i] * v[i];
      }
    }
  }
  block_flush_stats(acc, acc2, c, n, gsum, gsq, s_part);
}
// Separable [1,2,1]/4 x [1,2,1]/4 blur with a 3-row sliding window in registers: every thread owns 8 channels of one
// pixel column and walks down a band of kBlurRows rows, so each raw element is loaded ~3x from L1 and ~1x from
// HBM (the 9-tap form above loads it 9x).  Then +noise*w +bias, LeakyReLU(0.2), fp16 store, per-(n,c) sums.
// Strip height: 128 rows (2 halo rows re-read per 128 instead of per 32: L12 3.42 -> 3.1 us, L14 7.2 -> 6.5 us per sample;
// 256 / 512 are within the run-to-run noise of 128 and leave fewer blocks for the small layers).
#ifndef CFR_BLUR_ROWS
#define CFR_BLUR_ROWS 128
#endif
constexpr int kBlurRows = CFR_BLUR_ROWS;

struct Raw3 {
  uint4 l, c, r;
};
__device__ __forceinline__ Raw3 load_raw3(const __half* __restrict__ rowp, int px, int w, int c, int ch, bool row_ok) {
  Raw3 o;
  const uint4 z = make_uint4(0u, 0u, 0u, 0u);
  o.l = z; o.c = z; o.r = z;
  if (row_ok) {
    const __half* pc = rowp + static_cast<size_t>(px) * c + ch;
    o.c = __ldg(reinterpret_cast<const uint4*>(pc));
    if (px > 0) o.l = __ldg(reinterpret_cast<const uint4*>(pc - c));
    if (px < w - 1) o.r = __ldg(reinterpret_cast<const uint4*>(pc + c));
  }
  return o;
}
__device__ __forceinline__ void hblur8(const Raw3& t, float (&h)[8]) {
  const __half2* l2 = reinterpret_cast<const __half2*>(&t.l);
  const __half2* c2 = reinterpret_cast<const __half2*>(&t.c);
  const __half2* r2 = reinterpret_cast<const __half2*>(&t.r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = __half22float2(l2[i]), b = __half22float2(c2[i]), d = __half22float2(r2[i]);
    h[2 * i] = 0.25f * (a.x + d.x) + 0.5f * b.x;
    h[2 * i + 1] = 0.25f * (a.y + d.y) + 0.5f * b.y;
  }
}

__global__ void __launch_bounds__(256, 3) k_blur_rows(const __half* __restrict__ raw, __half* __restrict__ y, int h, int w,
                                                      int c, const float* __restrict__ noise,
                                                      const float* __restrict__ noise_w, const float* __restrict__ bias,
                                                      stat_t* __restrict__ gsum, stat_t* __restrict__ gsq) {
  __shared__ float s_part[2 * 2048];
  const int n = blockIdx.z;
  const int c8 = c >> 3;
  const int ppb = 256 / c8;
  const int cg = threadIdx.x % c8, pl = threadIdx.x / c8;
  const int ch = cg * 8;
  const int px = blockIdx.x * ppb + pl;
  const int y0 = blockIdx.y * kBlurRows;
  float acc[8], acc2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i] = 0.f;
    acc2[i] = 0.f;
  }
  if (px < w) {
    float nw[8], bs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      nw[i] = noise_w[ch + i];
      bs[i] = bias[ch + i];
    }
    const __half* img = raw + static_cast<size_t>(n) * h * w * c;
    const size_t rstride = static_cast<size_t>(w) * c;
    float hp[8], hc[8], hn[8];
    hblur8(load_raw3(img + (y0 - 1) * static_cast<long long>(rstride), px, w, c, ch, y0 - 1 >= 0), hp);
    hblur8(load_raw3(img + y0 * rstride, px, w, c, ch, true), hc);
    const int y1 = min(h, y0 + kBlurRows);
    // software pipeline: the three loads of row yy+2 are in flight while row yy is blurred, activated and stored
    Raw3 nxt = load_raw3(img + (y0 + 1) * rstride, px, w, c, ch, y0 + 1 < h);
    float nz = __ldg(&noise[y0 * w + px]);
    for (int yy = y0; yy < y1; ++yy) {
      const Raw3 cur = nxt;
      const float nzc = nz;
      if (yy + 1 < y1) {
        nxt = load_raw3(img + (yy + 2) * rstride, px, w, c, ch, yy + 2 < h);
        nz = __ldg(&noise[(yy + 1) * w + px]);
      }
      hblur8(cur, hn);
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float t = 0.25f * (hp[i] + hn[i]) + 0.5f * hc[i];
        t = fmaf(nzc, nw[i], t) + bs[i];
        t = fmaxf(t, 0.2f * t);
        v[i] = t;
        acc[i] += t;
        acc2[i] = fmaf(t, t, acc2[i]);
        hp[i] = hc[i];
        hc[i] = hn[i];
      }
      store8(y + ((static_cast<size_t>(n) * h + yy) * w + px) * c + ch, v);
    }
  }
  block_flush_stats(acc, acc2, c, n, gsum, gsq, s_part);
}

// cp.async-pipelined variant of k_blur_rows: the loads of the next kBlurStages-1 rows of a block's strip are in flight
// as 16-byte LDGSTS into a shared-memory ring (zero-filled outside the image == the blur's zero padding), so every
// raw element is fetched from L2/HBM once.  The noise row rides in the same ring.  ncu (profiles/ncu_r01_notes.md)
// showed the first version issue-bound (IPC 3.1, 274 SASS instructions per row of which ~140 were address
// arithmetic and register moves), hence: running pointers, the ring/role rotation unrolled 6x so stage offsets and
// the three-row window are compile-time, and packed fp32x2 arithmetic (sm_100 FADD2/FFMA2/FMUL2).
constexpr int kBlurStages = 6;
__device__ __forceinline__ f2_t f2_from_half2(uint32_t h2) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h2));
  return f2_pack(f.x, f.y);
}

__global__ void __launch_bounds__(256, 2) k_blur_pipe(const __half* __restrict__ raw, __half* __restrict__ y, int h, int w,
                                                      int c, const float* __restrict__ noise,
                                                      const float* __restrict__ noise_w, const float* __restrict__ bias,
                                                      stat_t* __restrict__ gsum, stat_t* __restrict__ gsq) {
  extern __shared__ uint4 ring[];                       // [stage][(ppb + 2) * c8] uint4, then [stage][ppb] noise floats
  const int n = blockIdx.z;
  const int c8 = c >> 3;
  const int ppb = 256 / c8;
  const int cg = threadIdx.x % c8, pl = threadIdx.x / c8;
  const int ch = cg * 8;
  const int px0 = blockIdx.x * ppb, px = px0 + pl;
  const int y0 = blockIdx.y * kBlurRows, y1 = min(h, y0 + kBlurRows);
  const uint32_t stage_bytes = static_cast<uint32_t>((ppb + 2) * c8) * 16u;
  const uint32_t ring_base = smem_u32(ring);
  const uint32_t nring_base = ring_base + kBlurStages * stage_bytes;
  const __half* img = raw + static_cast<size_t>(n) * h * w * c;
  const long long rstride_b = static_cast<long long>(w) * c * 2;          // bytes per image row
  const bool colok = px < w;
  // halo pixels (left of the first / right of the last pixel of the block) are fetched by the first 2*c8 threads
  const bool is_halo = threadIdx.x < 2 * c8;
  const int hside = threadIdx.x / c8;                   // 0 left, 1 right (only meaningful if is_halo)
  const int hpx = hside == 0 ? px0 - 1 : px0 + ppb;
  const bool hcolok = is_halo && hpx >= 0 && hpx < w;
  // running state of the producer side: raw row r_next goes to the stage the unrolled loop names statically
  int r_next = y0 - 1;
  const char* src = reinterpret_cast<const char*>(img) + r_next * rstride_b + (static_cast<long long>(px) * c + ch) * 2;
  const char* hsrc = reinterpret_cast<const char*>(img) + r_next * rstride_b +
                     (static_cast<long long>(hpx) * c + (threadIdx.x % c8) * 8) * 2;
  const float* nsrc = noise + static_cast<long long>(r_next - 1) * w + px;  // noise of output row r_next - 1
  const uint32_t own_dst = ring_base + static_cast<uint32_t>((pl + 1) * c8 + cg) * 16u;
  const uint32_t halo_dst = ring_base + static_cast<uint32_t>((hside == 0 ? 0 : ppb + 1) * c8 + (threadIdx.x % c8)) * 16u;
  const uint32_t noise_dst = nring_base + static_cast<uint32_t>(pl) * 4u;
  const bool noise_thr = cg == 0 && colok;
  const int r_last = min(y1, h - 1);                    // last raw row this strip needs that lies inside the image
  auto issue = [&](int st) {                            // raw row r_next (+ noise of output row r_next - 1) -> stage st
    const bool rok = r_next >= 0 && r_next <= r_last;
    cp_async16(own_dst + st * stage_bytes, (rok && colok) ? src : reinterpret_cast<const char*>(img), (rok && colok) ? 16u : 0u);
    if (is_halo)
      cp_async16(halo_dst + st * stage_bytes, (rok && hcolok) ? hsrc : reinterpret_cast<const char*>(img), (rok && hcolok) ? 16u : 0u);
    if (noise_thr && r_next > y0 && r_next <= y1)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(noise_dst + st * (ppb * 4)), "l"(nsrc) : "memory");
    cp_async_commit();
    ++r_next;
    src += rstride_b;
    hsrc += rstride_b;
    nsrc += w;
  };
#pragma unroll
  for (int i = 0; i < kBlurStages - 1; ++i) issue(i);

  f2_t nw2[4], bs2[4], acc[4], acc2[4], H[3][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    // threads right of the image (colok false) compute with zero gains: their u is 0, so the sums need no masking
    nw2[i] = colok ? f2_pack(noise_w[ch + 2 * i], noise_w[ch + 2 * i + 1]) : f2_pack(0.f, 0.f);
    bs2[i] = colok ? f2_pack(bias[ch + 2 * i], bias[ch + 2 * i + 1]) : f2_pack(0.f, 0.f);
    acc[i] = f2_pack(0.f, 0.f);
    acc2[i] = acc[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) H[j][i] = acc[i];
  }
  const float sx = colok ? 0.0625f : 0.f;
  const f2_t two2 = f2_pack(2.f, 2.f), sixteenth2 = f2_pack(sx, sx), slope2 = f2_pack(0.2f, 0.2f);
  const uint32_t rd_base = ring_base + static_cast<uint32_t>(pl * c8 + cg) * 16u;
  const uint32_t nrd_base = nring_base + static_cast<uint32_t>(pl) * 4u;
  __half* outp = y + ((static_cast<size_t>(n) * h + y0) * w + (colok ? px : 0)) * c + ch;   // output row y0 + (k - 2)
  const size_t ostride = static_cast<size_t>(w) * c;
  // one row step: wait for raw row k (stage J), refill the stage freed by the previous step, horizontal blur of row k
  // into H[J % 3]; if OUT, emit output row k - 2 from the three-row window
  auto step = [&](auto Jc, auto OUTc) {
    constexpr int J = decltype(Jc)::value;
    constexpr bool OUT = decltype(OUTc)::value;
    cp_async_wait<kBlurStages - 2>();
    __syncthreads();
    issue((J + kBlurStages - 1) % kBlurStages);
    uint32_t lw[4], mw[4], rw[4];
    const uint32_t a0 = rd_base + J * stage_bytes;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(lw[0]), "=r"(lw[1]), "=r"(lw[2]), "=r"(lw[3]) : "r"(a0));
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(mw[0]), "=r"(mw[1]), "=r"(mw[2]), "=r"(mw[3]) : "r"(a0 + c8 * 16));
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]) : "r"(a0 + c8 * 32));
    f2_t(&hn)[4] = H[J % 3];
    f2_t(&hc)[4] = H[(J + 2) % 3];
    f2_t(&hp)[4] = H[(J + 1) % 3];
#pragma unroll
    for (int i = 0; i < 4; ++i)                         // horizontal blur x4: (l + r) + 2 m
      hn[i] = f2_fma(two2, f2_from_half2(mw[i]), f2_add(f2_from_half2(lw[i]), f2_from_half2(rw[i])));
    if constexpr (OUT) {
      float nz;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(nz) : "r"(nrd_base + J * (ppb * 4)));
      const f2_t nz2 = f2_pack(nz, nz);
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const f2_t vs = f2_fma(two2, hc[i], f2_add(hp[i], hn[i]));              // 16 x blurred value
        const f2_t t = f2_fma(vs, sixteenth2, f2_fma(nz2, nw2[i], bs2[i]));
        const float2 tf = f2_unpack(t), sf = f2_unpack(f2_mul(t, slope2));
        const float u0 = fmaxf(tf.x, sf.x), u1 = fmaxf(tf.y, sf.y);
        const f2_t u = f2_pack(u0, u1);
        acc[i] = f2_add(acc[i], u);
        acc2[i] = f2_fma(u, u, acc2[i]);
        const __half2 hh = __floats2half2_rn(u0, u1);
        o[i] = *reinterpret_cast<const uint32_t*>(&hh);
      }
      if (colok) *reinterpret_cast<uint4*>(outp) = make_uint4(o[0], o[1], o[2], o[3]);
      outp += ostride;
    }
  };
  using std::integral_constant;
  step(integral_constant<int, 0>{}, std::false_type{});          // raw rows y0 - 1 and y0 only fill the window
  step(integral_constant<int, 1>{}, std::false_type{});
  const int rows = y1 - y0 + 2;
  for (int k0 = 2; k0 < rows; k0 += kBlurStages) {              // `k0 + j < rows` is block-uniform
    step(integral_constant<int, 2>{}, std::true_type{});
    if (k0 + 1 < rows) step(integral_constant<int, 3>{}, std::true_type{});
    if (k0 + 2 < rows) step(integral_constant<int, 4>{}, std::true_type{});
    if (k0 + 3 < rows) step(integral_constant<int, 5>{}, std::true_type{});
    if (k0 + 4 < rows) step(integral_constant<int, 0>{}, std::true_type{});
    if (k0 + 5 < rows) step(integral_constant<int, 1>{}, std::true_type{});
  }
  cp_async_wait<0>();
  __syncthreads();                                       // the ring is dead: reuse it for the block reduction
  float facc[8], facc2[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = f2_unpack(acc[i]), b = f2_unpack(acc2[i]);
    facc[2 * i] = a.x; facc[2 * i + 1] = a.y;
    facc2[2 * i] = b.x; facc2[2 * i + 1] = b.y;
  }
  block_flush_stats(facc, facc2, c, n, gsum, gsq, reinterpret_cast<float*>(ring));
}

int launch_blur_act_stats(const __half* raw, __half* y, int n, int h, int w, int c, const float* noise,
                          const float* noise_w, const float* bias, void* sum_v, void* sq_v, int mode, cudaStream_t st) {
  stat_t* sum = static_cast<stat_t*>(sum_v);
  stat_t* sq = static_cast<stat_t*>(sq_v);
  if (c < 8 || c > 512 || (c & (c - 1)) != 0) { set_error("blur_act_stats: C=%d unsupported (power of two in 8..512)", c); return 2; }
  const int ppb = 256 / (c / 8);
  if (mode == 0) {
    dim3 grid((w + ppb - 1) / ppb, (h + kBlurRows - 1) / kBlurRows, n);
    static const bool legacy = getenv("CFR_BLUR_LEGACY") != nullptr;      // A/B knob for profiling
    if (legacy || c > 256) {             // c8 > 32 would need > 2*c8 halo threads per 256; those layers are tiny anyway
      k_blur_rows<<<grid, 256, 0, st>>>(raw, y, h, w, c, noise, noise_w, bias, sum, sq);
    } else {
      const size_t smem = static_cast<size_t>(kBlurStages) * ((ppb + 2) * (c / 8) * 16 + ppb * 4);
      k_blur_pipe<<<grid, 256, smem, st>>>(raw, y, h, w, c, noise, noise_w, bias, sum, sq);
    }
    CFR_LAUNCH_CHECK("blur_rows");
    return 0;
  }
  int bx = (h * w + ppb - 1) / ppb;
  const int cap = 16 * 148 / (n > 0 ? n : 1) + 1;       // ~16 blocks per SM in total
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n);
  k_blur_act_stats<1><<<grid, 256, 0, st>>>(raw, y, h, w, c, noise, noise_w, bias, sum, sq);
  CFR_LAUNCH_CHECK("blur_act_stats");
  return 0;
}

// fp32 in / fp32 out (split-precision early layers; at most 64x64 x 256 channels: the simple kernel is enough)
int launch_blur_act_stats_f32(const float* raw, float* y, int n, int h, int w, int c, const float* noise,
                              const float* noise_w, const float* bias, void* sum_v, void* sq_v, int mode, cudaStream_t st) {
  stat_t* sum = static_cast<stat_t*>(sum_v);
  stat_t* sq = static_cast<stat_t*>(sq_v);
  if (c < 8 || c > 512 || (c & (c - 1)) != 0) { set_error("blur_act_stats_f32: C=%d unsupported (power of two in 8..512)", c); return 2; }
  const int ppb = 256 / (c / 8);
  int bx = (h * w + ppb - 1) / ppb;
  const int cap = 16 * 148 / (n > 0 ? n : 1) + 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid(bx, n);
  if (mode == 0) k_blur_act_stats<0, float><<<grid, 256, 0, st>>>(raw, y, h, w, c, noise, noise_w, bias, sum, sq);
  else k_blur_act_stats<1, float><<<grid, 256, 0, st>>>(raw, y, h, w, c, noise, noise_w, bias, sum, sq);
  CFR_LAUNCH_CHECK("blur_act_stats_f32");
  return 0;
}

// ------------------------------------------------------------------------------------------
__global__ void k_finalize_stats(const long long* __restrict__ sum, const long long* __restrict__ sq,
                                 const float* __restrict__ styles, int style_stride, int style_off, int n, int c,
                                 float inv_count, float* __restrict__ A, float* __restrict__ B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int s = i / c, ch = i - s * c;
  const double meand = static_cast<double>(sum[i]) * (1.0 / kStatScale) * inv_count;
  const float mean = static_cast<float>(meand);
  float var = static_cast<float>(static_cast<double>(sq[i]) * (1.0 / kStatScale) * inv_count - meand * meand);
  var = var > 0.f ? var : 0.f;
  const float rstd = 1.0f / sqrtf(var + 1e-8f);
  const float s0 = styles[static_cast<size_t>(s) * style_stride + style_off + ch];
  const float s1 = styles[static_cast<size_t>(s) * style_stride + style_off + c + ch];
  const float a = rstd * (s0 + 1.f);
  A[i] = a;
  B[i] = s1 - mean * a;
}
int launch_finalize_stats(const void* sum, const void* sq, const float* styles, int style_stride, int style_off,
                          int n, int c, float inv_count, float* A, float* B, cudaStream_t st) {
  k_finalize_stats<<<(n * c + 255) / 256, 256, 0, st>>>(static_cast<const long long*>(sum), static_cast<const long long*>(sq), styles, style_stride, style_off, n, c, inv_count, A, B);
  CFR_LAUNCH_CHECK("finalize_stats");
  return 0;
}

__global__ void k_affine(const __half* __restrict__ y, const float* __restrict__ A, const float* __restrict__ B,
                         size_t total8, int hw, int c, __half* __restrict__ x) {
  const int c8 = c >> 3;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    const size_t pix = i / c8;
    const int n = static_cast<int>(pix / hw);
    float v[8];
    load8(y + i * 8, v);
    const float4* a4 = reinterpret_cast<const float4*>(A + static_cast<size_t>(n) * c + cg * 8);
    const float4* b4 = reinterpret_cast<const float4*>(B + static_cast<size_t>(n) * c + cg * 8);
    const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1), b0 = __ldg(b4), b1 = __ldg(b4 + 1);
    v[0] = v[0] * a0.x + b0.x; v[1] = v[1] * a0.y + b0.y; v[2] = v[2] * a0.z + b0.z; v[3] = v[3] * a0.w + b0.w;
    v[4] = v[4] * a1.x + b1.x; v[5] = v[5] * a1.y + b1.y; v[6] = v[6] * a1.z + b1.z; v[7] = v[7] * a1.w + b1.w;
    store8(x + i * 8, v);
  }
}
int launch_affine(const __half* y, const float* A, const float* B, int n, int hw, int c, __half* x, cudaStream_t st) {
  const size_t total8 = static_cast<size_t>(n) * hw * c / 8;
  size_t blocks = (total8 + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  k_affine<<<static_cast<unsigned>(blocks), 256, 0, st>>>(y, A, B, total8, hw, c, x);
  CFR_LAUNCH_CHECK("affine");
  return 0;
}

// fp32 y -> x = y*A + B as fp16: SPLIT == 1 plain [.., c]; SPLIT == 3 [.., 3c] = [hi | lo | hi], lo = fp16(x - hi)
template <int SPLIT>
__global__ void k_affine_f32(const float* __restrict__ y, const float* __restrict__ A, const float* __restrict__ B,
                             size_t total8, int hw, int c, __half* __restrict__ x) {
  const int c8 = c >> 3;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total8;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    const size_t pix = i / c8;
    const int n = static_cast<int>(pix / hw);
    float v[8];
    load8(y + i * 8, v);
    const float4* a4 = reinterpret_cast<const float4*>(A + static_cast<size_t>(n) * c + cg * 8);
    const float4* b4 = reinterpret_cast<const float4*>(B + static_cast<size_t>(n) * c + cg * 8);
    const float4 a0 = __ldg(a4), a1 = __ldg(a4 + 1), b0 = __ldg(b4), b1 = __ldg(b4 + 1);
    v[0] = v[0] * a0.x + b0.x; v[1] = v[1] * a0.y + b0.y; v[2] = v[2] * a0.z + b0.z; v[3] = v[3] * a0.w + b0.w;
    v[4] = v[4] * a1.x + b1.x; v[5] = v[5] * a1.y + b1.y; v[6] = v[6] * a1.z + b1.z; v[7] = v[7] * a1.w + b1.w;
    if (SPLIT == 1) {
      store8(x + i * 8, v);
    } else {
      float lo[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float hi = __half2float(__float2half_rn(v[k]));
        lo[k] = v[k] - hi;
        v[k] = hi;
      }
      __half* o = x + pix * (3 * static_cast<size_t>(c)) + cg * 8;
      store8(o, v);
      store8(o + c, lo);
      store8(o + 2 * c, v);
    }
  }
}
int launch_affine_f32(const float* y, const float* A, const float* B, int n, int hw, int c, __half* x, int split,
                      cudaStream_t st) {
  if (split != 1 && split != 3) { set_error("affine_f32: split must be 1 or 3"); return 2; }
  const size_t total8 = static_cast<size_t>(n) * hw * c / 8;
  size_t blocks = (total8 + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (split == 3) k_affine_f32<3><<<static_cast<unsigned>(blocks), 256, 0, st>>>(y, A, B, total8, hw, c, x);
  else k_affine_f32<1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(y, A, B, total8, hw, c, x);
  CFR_LAUNCH_CHECK("affine_f32");
  return 0;
}

// ------------------------------------------------------------------------------------------
// toRGB (1x1, C->3) + postprocess + bilinear (align_corners=False, no antialias) + normalise.
// Only the 4 source pixels each output pixel needs are read: 4 x C fp16 per output pixel.
// ------------------------------------------------------------------------------------------
__global__ void k_torgb_resize(const __half* __restrict__ x, const float* __restrict__ A, const float* __restrict__ B,
                               int n, int hin, int c, const float* __restrict__ w_rgb, const float* __restrict__ b_rgb,
                               int rout, float mean, float stdv, __half* __restrict__ out,
                               float* __restrict__ out_planar, const int* __restrict__ slot,
                               const int* __restrict__ keep_map, int keep_dim) {
  // keep_map != nullptr: x is the compact [n][keep_dim][keep_dim][c] output of a sparse-store conv; source row / column
  // y lives at index keep_map[y] (the conv kept exactly the rows / columns this resize reads; engine.resize_keep_map)
  // one sample per blockIdx.y: toRGB weights with the sample's IN/AdaIN folded in live in shared memory
  //   rgb_k = sum_c w[k][c] * (A_c x_c + B_c) + b_k = sum_c (w[k][c] A_c) x_c + (b_k + sum_c w[k][c] B_c)
  extern __shared__ float trs[];                 // [3][c] fused weights, then [3] fused bias
  const int s = blockIdx.y;
  for (int t = threadIdx.x; t < 3 * c; t += blockDim.x)
    trs[t] = w_rgb[t] * (A != nullptr ? A[s * c + t % c] : 1.f);
  if (threadIdx.x < 3) {
    float bsum = b_rgb[threadIdx.x];
    if (B != nullptr)
      for (int cc = 0; cc < c; ++cc) bsum = fmaf(w_rgb[threadIdx.x * c + cc], B[s * c + cc], bsum);
    trs[3 * c + threadIdx.x] = bsum;
  }
  __syncthreads();
  const int ip = blockIdx.x * blockDim.x + threadIdx.x;      // pixel inside the sample
  if (ip >= rout * rout) return;
  const int i = s * rout * rout + ip;
  if (slot != nullptr) {                    // write into the slot-th group of n images of a larger buffer
    const size_t g = static_cast<size_t>(*slot) * n * rout * rout;
    if (out != nullptr) out += g * 16;
    if (out_planar != nullptr) out_planar += g * 3;
  }
  const int ox = ip % rout, oy = ip / rout;
  const float scale = static_cast<float>(hin) / static_cast<float>(rout);
  float sy = scale * (oy + 0.5f) - 0.5f;
  float sx = scale * (ox + 0.5f) - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
  const int y1 = y0 + (y0 < hin - 1 ? 1 : 0), x1 = x0 + (x0 < hin - 1 ? 1 : 0);
  const float ly1 = sy - y0, lx1 = sx - x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const int ys[2] = {y0, y1}, xs[2] = {x0, x1};
  float px[2][2][3];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
#pragma unroll
    for (int bq = 0; bq < 2; ++bq) {
      float r = 0.f, g = 0.f, bl = 0.f;
      const __half* src = keep_map != nullptr
          ? x + ((static_cast<size_t>(s) * keep_dim + __ldg(keep_map + ys[a])) * keep_dim + __ldg(keep_map + xs[bq])) * c
          : x + ((static_cast<size_t>(s) * hin + ys[a]) * hin + xs[bq]) * c;
      for (int cc = 0; cc < c; cc += 8) {
        float v[8];
        load8(src + cc, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          r = fmaf(trs[cc + k], v[k], r);
          g = fmaf(trs[c + cc + k], v[k], g);
          bl = fmaf(trs[2 * c + cc + k], v[k], bl);
        }
      }
      const float rgb[3] = {r + trs[3 * c], g + trs[3 * c + 1], bl + trs[3 * c + 2]};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float t = (rgb[k] + 1.0f) / 2.0f + 0.5f / 255.f;
        px[a][bq][k] = fminf(fmaxf(t, 0.f), 1.f);
      }
    }
  }
  float res[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float v = ly0 * (lx0 * px[0][0][k] + lx1 * px[0][1][k]) + ly1 * (lx0 * px[1][0][k] + lx1 * px[1][1][k]);
    res[k] = (v - mean) / stdv;
  }
  if (out != nullptr) {
    float o[8] = {res[0], res[1], res[2], 0.f, 0.f, 0.f, 0.f, 0.f};
    const float zz[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store8(out + static_cast<size_t>(i) * 16, o);
    store8(out + static_cast<size_t>(i) * 16 + 8, zz);
  }
  if (out_planar != nullptr) {
#pragma unroll
    for (int k = 0; k < 3; ++k) out_planar[((static_cast<size_t>(s) * 3 + k) * rout + oy) * rout + ox] = res[k];
  }
}
int launch_torgb_resize(const __half* x, const float* A, const float* B, int n, int hin, int c, const float* w_rgb,
                        const float* b_rgb, int rout, float mean, float stdv, __half* out, float* out_planar,
                        const int* slot, cudaStream_t st, const int* keep_map, int keep_dim) {
  if (c > 2048) { set_error("torgb_resize: c=%d too large", c); return 2; }
  k_torgb_resize<<<dim3((rout * rout + 127) / 128, n), 128, (3 * c + 3) * sizeof(float), st>>>(x, A, B, n, hin, c, w_rgb, b_rgb, rout, mean, stdv, out, out_planar, slot, keep_map, keep_dim);
  CFR_LAUNCH_CHECK("torgb_resize");
  return 0;
}

// ------------------------------------------------------------------------------------------
// gallery match: key = (float bits of squared distance) << 32 | row  ->  atomicMin == argmin with
// first-index tie-break (torch.argmax semantics on softmax(-d)).
// ------------------------------------------------------------------------------------------
constexpr int kMatchEmb = 8;      // embeddings per block
constexpr int kMatchRows = 64;    // gallery rows per block

__global__ void __launch_bounds__(256) k_match(const float* __restrict__ emb, int b, const float* __restrict__ gallery,
                                               int n, unsigned long long* __restrict__ keys, unsigned row0) {
  __shared__ float4 s_emb[kMatchEmb][128];
  const int e0 = blockIdx.y * kMatchEmb;
  for (int i = threadIdx.x; i < kMatchEmb * 128; i += 256) {
    const int e = i / 128, k = i % 128;
    s_emb[e][k] = (e0 + e < b) ? reinterpret_cast<const float4*>(emb + static_cast<size_t>(e0 + e) * 512)[k]
                               : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned long long best[kMatchEmb];
#pragma unroll
  for (int e = 0; e < kMatchEmb; ++e) best[e] = ~0ull;
  const int r_end = min(n, (blockIdx.x + 1) * kMatchRows);
  for (int r = blockIdx.x * kMatchRows + warp; r < r_end; r += 8) {
    float4 g[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) g[k] = __ldg(reinterpret_cast<const float4*>(gallery + static_cast<size_t>(r) * 512) + k * 32 + lane);
#pragma unroll
    for (int e = 0; e < kMatchEmb; ++e) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 v = s_emb[e][k * 32 + lane];
        const float dx = v.x - g[k].x, dy = v.y - g[k].y, dz = v.z - g[k].z, dw = v.w - g[k].w;
        acc += dx * dx + dy * dy + dz * dz + dw * dw;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      const unsigned long long key = (static_cast<unsigned long long>(__float_as_uint(acc)) << 32) | (static_cast<unsigned>(r) + row0);
      best[e] = key < best[e] ? key : best[e];
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int e = 0; e < kMatchEmb; ++e)
      if (e0 + e < b && best[e] != ~0ull) atomicMin(&keys[e0 + e], best[e]);
  }
}
__global__ void k_vote(unsigned long long* __restrict__ keys, int b, int n, int* __restrict__ pred,
                       unsigned long long* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  const int pr = static_cast<int>(keys[i] & 0xffffffffull);
  keys[i] = ~0ull;                 // re-arm for the next batch
  if (pred != nullptr) pred[i] = pr;
  // (a key nobody wrote decodes to row 0xffffffff: never index the tally with it)
  if (counts != nullptr && static_cast<unsigned>(pr) < static_cast<unsigned>(n)) atomicAdd(&counts[pr], 1ull);
}
// ------------------------------------------------------------------------------------------
// Tensor-core gallery match (large galleries): fp32 rows are split into fp16 hi + lo parts so that
//   e.g = e_h.g_h + e_h.g_l + e_l.g_h  (+ O(2^-22)),  one K = 3*512 GEMM:  A = [e_h, e_h, e_l],  B = [g_h, g_l, g_h].
// mode 0 (gallery): dst = [h, l, h], bias[row] = -|g|^2 (exact fp32; padded rows get -inf);
// mode 1 (queries): dst = [2h, 2h, 2l]  (the factor 2 of 2 e.g - |g|^2 is exact in fp16).
// ------------------------------------------------------------------------------------------
__global__ void k_split_hilo(const float* __restrict__ src, int rows, int rows_pad, int mode, __half* __restrict__ dst,
                             float* __restrict__ bias) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows_pad) return;
  __half* d = dst + static_cast<size_t>(row) * 1536;
  float nrm = 0.f;
  for (int k = lane; k < 512; k += 32) {
    const float v = row < rows ? src[static_cast<size_t>(row) * 512 + k] : 0.f;
    const float s = mode == 1 ? 2.f * v : v;
    const __half h = __float2half_rn(s);
    const __half l = __float2half_rn(s - __half2float(h));
    d[k] = h;
    d[512 + k] = mode == 1 ? h : l;
    d[1024 + k] = mode == 1 ? l : h;
    nrm += v * v;
  }
  if (bias != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, o);
    if (lane == 0) bias[row] = row < rows ? -nrm : -INFINITY;
  }
}
int launch_split_hilo(const float* src, int rows, int rows_pad, int mode, __half* dst, float* bias, cudaStream_t st) {
  k_split_hilo<<<(rows_pad + 7) / 8, 256, 0, st>>>(src, rows, rows_pad, mode, dst, bias);
  CFR_LAUNCH_CHECK("split_hilo");
  return 0;
}
__global__ void k_vote_argmax(unsigned long long* __restrict__ keys, int b, int n, int* __restrict__ pred,
                              unsigned long long* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  const int pr = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(keys[i] & 0xffffffffull));
  keys[i] = 0ull;                  // re-arm (atomicMax identity)
  if (pred != nullptr) pred[i] = pr;
  if (counts != nullptr && static_cast<unsigned>(pr) < static_cast<unsigned>(n)) atomicAdd(&counts[pr], 1ull);
}
int launch_vote_argmax(unsigned long long* keys, int b, int n, int* pred, long long* counts, cudaStream_t st) {
  if (b <= 0) return 0;
  k_vote_argmax<<<(b + 127) / 128, 128, 0, st>>>(keys, b, n, pred, reinterpret_cast<unsigned long long*>(counts));
  CFR_LAUNCH_CHECK("vote_argmax");
  return 0;
}

__global__ void k_set_int(int* p, int v) { *p = v; }
int launch_set_int(int* p, int v, cudaStream_t st) {
  k_set_int<<<1, 1, 0, st>>>(p, v);
  CFR_LAUNCH_CHECK("set_int");
  return 0;
}

int launch_match_vote(const float* emb, int b, const float* gallery, int n, unsigned long long* keys, int* pred,
                      long long* counts, cudaStream_t st) {
  if (b <= 0) return 0;
  dim3 grid((n + kMatchRows - 1) / kMatchRows, (b + kMatchEmb - 1) / kMatchEmb);
  k_match<<<grid, 256, 0, st>>>(emb, b, gallery, n, keys, 0u);
  CFR_LAUNCH_CHECK("match");
  k_vote<<<(b + 127) / 128, 128, 0, st>>>(keys, b, n, pred, reinterpret_cast<unsigned long long*>(counts));
  CFR_LAUNCH_CHECK("vote");
  return 0;
}

// out[n][c] = bias[c] + sum_p in[n][p][c] in a fixed order (split-K partial sums of the ArcFace FC)
__global__ void k_sum_partials(const float* __restrict__ in, const float* __restrict__ bias, int n, int parts, int c,
                               float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * c) return;
  const int s = i / c, ch = i - s * c;
  float a = in[(static_cast<size_t>(s) * parts) * c + ch];
  for (int p = 1; p < parts; ++p) a += in[(static_cast<size_t>(s) * parts + p) * c + ch];
  out[i] = a + (bias != nullptr ? bias[ch] : 0.f);
}
int launch_sum_partials(const float* in, const float* bias, int n, int parts, int c, float* out, cudaStream_t st) {
  if (n <= 0) return 0;
  k_sum_partials<<<(n * c + 255) / 256, 256, 0, st>>>(in, bias, n, parts, c, out);
  CFR_LAUNCH_CHECK("sum_partials");
  return 0;
}

// ---- InceptionResnetV1 glue kernels (tiny, HBM-trivial) ------------------------------------------------------
__global__ void k_maxpool3s2(const __half* __restrict__ in, int n, int h, int w, int c, __half* __restrict__ out,
                             int out_c_total, int c_off) {
  const int ho = (h - 3) / 2 + 1, wo = (w - 3) / 2 + 1, c8 = c >> 3;
  const size_t total = static_cast<size_t>(n) * ho * wo * c8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(i % c8);
    size_t r = i / c8;
    const int ox = static_cast<int>(r % wo);
    r /= wo;
    const int oy = static_cast<int>(r % ho), s = static_cast<int>(r / ho);
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -3.0e38f;
    for (int dy = 0; dy < 3; ++dy)
      for (int dx = 0; dx < 3; ++dx) {
        float v[8];
        load8(in + ((static_cast<size_t>(s) * h + 2 * oy + dy) * w + 2 * ox + dx) * c + cg * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], v[k]);
      }
    store8(out + ((static_cast<size_t>(s) * ho + oy) * wo + ox) * out_c_total + c_off + cg * 8, m);
  }
}
int launch_maxpool3s2(const __half* in, int n, int h, int w, int c, __half* out, int out_c_total, int c_off, cudaStream_t st) {
  if (c % 8 != 0 || c_off % 8 != 0 || out_c_total % 8 != 0 || h < 3 || w < 3) { set_error("maxpool3s2: bad shape"); return 2; }
  const size_t total = static_cast<size_t>(n) * ((h - 3) / 2 + 1) * ((w - 3) / 2 + 1) * (c / 8);
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_maxpool3s2<<<static_cast<unsigned>(blocks), 256, 0, st>>>(in, n, h, w, c, out, out_c_total, c_off);
  CFR_LAUNCH_CHECK("maxpool3s2");
  return 0;
}
__global__ void k_avgpool(const __half* __restrict__ in, int n, int hw, int c, __half* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;            // over n * c/8
  const int c8 = c >> 3;
  if (i >= n * c8) return;
  const int cg = i % c8, s = i / c8;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = 0; p < hw; ++p) {
    float v[8];
    load8(in + (static_cast<size_t>(s) * hw + p) * c + cg * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += v[k];
  }
  const float inv = 1.f / static_cast<float>(hw);
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] *= inv;
  store8(out + static_cast<size_t>(s) * c + cg * 8, a);
}
int launch_avgpool(const __half* in, int n, int hw, int c, __half* out, cudaStream_t st) {
  if (c % 8 != 0) { set_error("avgpool: C=%d must be a multiple of 8", c); return 2; }
  k_avgpool<<<(n * (c / 8) + 127) / 128, 128, 0, st>>>(in, n, hw, c, out);
  CFR_LAUNCH_CHECK("avgpool");
  return 0;
}
__global__ void k_l2norm(const float* __restrict__ in, int n, int c, float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  float q = 0.f;
  for (int k = lane; k < c; k += 32) {
    const float v = in[static_cast<size_t>(row) * c + k];
    q = fmaf(v, v, q);
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float inv = 1.f / fmaxf(sqrtf(q), 1e-12f);                 // F.normalize: x / max(||x||, eps)
  for (int k = lane; k < c; k += 32) out[static_cast<size_t>(row) * c + k] = in[static_cast<size_t>(row) * c + k] * inv;
}
int launch_l2norm(const float* in, int n, int c, float* out, cudaStream_t st) {
  k_l2norm<<<(n + 7) / 8, 256, 0, st>>>(in, n, c, out);
  CFR_LAUNCH_CHECK("l2norm");
  return 0;
}

// ---- gallery sharded over ranks (partition C) ----------------------------------------
// Every rank reduces its own rows to one 64-bit key per query such that the UNSIGNED MINIMUM over ranks is the global
// winner with torch.argmax's first-index tie-break:  exact matcher: (distance bits << 32) | global row;  tensor-core
// matcher: ~((ordered score bits << 32) | (0xFFFFFFFF - global row)).  The low word is always the global row.
int launch_match_keys(const float* emb, int b, const float* gallery, int n, unsigned row0, unsigned long long* keys,
                      cudaStream_t st) {
  if (b <= 0) return 0;
  cudaError_t e = cudaMemsetAsync(keys, 0xFF, sizeof(unsigned long long) * b, st);
  if (e != cudaSuccess) { set_error("match_keys memset: %s", cudaGetErrorString(e)); return 5; }
  dim3 grid((n + kMatchRows - 1) / kMatchRows, (b + kMatchEmb - 1) / kMatchEmb);
  k_match<<<grid, 256, 0, st>>>(emb, b, gallery, n, keys, row0);
  CFR_LAUNCH_CHECK("match_keys");
  return 0;
}
__global__ void k_export_argmax_keys(unsigned long long* __restrict__ keys, int b, unsigned row0,
                                     unsigned long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  const unsigned long long k = keys[i];
  keys[i] = 0ull;                  // re-arm (atomicMax identity)
  const unsigned col = 0xFFFFFFFFu - static_cast<unsigned>(k & 0xffffffffull);
  out[i] = ~((k & 0xffffffff00000000ull) | (0xFFFFFFFFu - (col + row0)));
}
int launch_export_argmax_keys(unsigned long long* keys, int b, unsigned row0, unsigned long long* out, cudaStream_t st) {
  if (b <= 0) return 0;
  k_export_argmax_keys<<<(b + 127) / 128, 128, 0, st>>>(keys, b, row0, out);
  CFR_LAUNCH_CHECK("export_argmax_keys");
  return 0;
}
__global__ void k_vote_keys(const unsigned long long* __restrict__ keys, int b, int* __restrict__ pred,
                            unsigned long long* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= b) return;
  const int pr = static_cast<int>(keys[i] & 0xffffffffull);
  if (pred != nullptr) pred[i] = pr;
  if (counts != nullptr) atomicAdd(&counts[pr], 1ull);
}
int launch_vote_keys(const unsigned long long* keys, int b, int* pred, long long* counts, cudaStream_t st) {
  if (b <= 0) return 0;
  k_vote_keys<<<(b + 127) / 128, 128, 0, st>>>(keys, b, pred, reinterpret_cast<unsigned long long*>(counts));
  CFR_LAUNCH_CHECK("vote_keys");
  return 0;
}

}  // namespace cfr
