// ATen-compatible bicubic (A = -0.75, clamped taps) and bilinear upsampling, align_corners=False,
// as reached from MyEfficientLFNet.py:88-90, EPIT.py:164-169 and DistgSSR.py:30 via F.interpolate.
// (utils/imresize.py is a different, MATLAB-style bicubic used only offline - SURVEY 2.2.)
// One thread per output pixel; HBM-bound on the s*s-times larger output, input rows are L1/L2 hits.
#include "lfsr_common.cuh"

namespace lfsr {

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__device__ __forceinline__ void cubic_coeffs(float t, float c[4]) {
  const float A = -0.75f;
  c[0] = cubic2(t + 1.f, A);
  c[1] = cubic1(t, A);
  c[2] = cubic1(1.f - t, A);
  c[3] = cubic2(2.f - t, A);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void __launch_bounds__(256)
interp_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int h, int w, int s, int mode, int bh,
              int bw) {
  const int oh = h * s, ow = w * s;
  const float rs = 1.0f / (float)s;
  const int img = blockIdx.y;                 // 32-bit index math inside one image
  out += (size_t)img * oh * ow;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < oh * ow; t += gridDim.x * blockDim.x) {
    const int ox = t % ow;
    const int oy = t / ow;
    // independent block (view) handling: clamp inside the block the output pixel falls in
    int vby = oy / (bh * s), ly = oy - vby * bh * s;
    int vbx = ox / (bw * s), lx = ox - vbx * bw * s;
    const float* base = in + (size_t)img * h * w + (size_t)vby * bh * w + (size_t)vbx * bw;
    float val;
    if (mode == LFSR_INTERP_BICUBIC) {
      float ry = rs * ((float)ly + 0.5f) - 0.5f;
      float rx = rs * ((float)lx + 0.5f) - 0.5f;
      float fy = floorf(ry), fx = floorf(rx);
      int iy = (int)fy, ix = (int)fx;
      float cy[4], cx[4];
      cubic_coeffs(ry - fy, cy);
      cubic_coeffs(rx - fx, cx);
      int xs[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xs[i] = clampi(ix - 1 + i, 0, bw - 1);
      val = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* row = base + (size_t)clampi(iy - 1 + j, 0, bh - 1) * w;
        float rsum = __ldg(row + xs[0]) * cx[0] + __ldg(row + xs[1]) * cx[1] + __ldg(row + xs[2]) * cx[2] +
                     __ldg(row + xs[3]) * cx[3];
        val += rsum * cy[j];
      }
    } else {
      float ry = fmaxf(rs * ((float)ly + 0.5f) - 0.5f, 0.f);
      float rx = fmaxf(rs * ((float)lx + 0.5f) - 0.5f, 0.f);
      int y0 = (int)ry, x0 = (int)rx;
      int y1 = y0 + (y0 < bh - 1 ? 1 : 0), x1 = x0 + (x0 < bw - 1 ? 1 : 0);
      float ly1 = ry - (float)y0, lx1 = rx - (float)x0;
      float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
      const float* r0 = base + (size_t)y0 * w;
      const float* r1 = base + (size_t)y1 * w;
      val = ly0 * (lx0 * __ldg(r0 + x0) + lx1 * __ldg(r0 + x1)) + ly1 * (lx0 * __ldg(r1 + x0) + lx1 * __ldg(r1 + x1));
    }
    out[t] = val;
  }
}

// Bicubic, scale S in {2, 4}: one thread per output row and INPUT column k writes the S outputs S*k .. S*k+S-1. Their taps
// all lie in input columns k-2 .. k+2, so a thread reads 5 x 4 values for S outputs (16 per output in the generic kernel),
// does its index arithmetic once and stores one 8/16-byte vector. Same expression per output as interp_kernel.
template <int S>
__global__ void __launch_bounds__(256)
interp_bicubic_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w, int bh, int bw) {
  const int oh = h * S, ow = w * S;
  const float rs = 1.0f / (float)S;
  const int img = blockIdx.y;
  out += (size_t)img * oh * ow;
  in += (size_t)img * h * w;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < oh * w; t += gridDim.x * blockDim.x) {
    const int kx = t % w;                      // input column
    const int oy = t / w;
    const int vby = oy / (bh * S), ly = oy - vby * bh * S;
    const int vbx = kx / bw, k = kx - vbx * bw;
    const float* base = in + (size_t)vby * bh * w + (size_t)vbx * bw;
    const float ry = rs * ((float)ly + 0.5f) - 0.5f;
    const float fy = floorf(ry);
    const int iy = (int)fy;
    float cy[4];
    cubic_coeffs(ry - fy, cy);
    float v[4][5];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float* row = base + (size_t)clampi(iy - 1 + j, 0, bh - 1) * w;
#pragma unroll
      for (int i = 0; i < 5; ++i) v[j][i] = __ldg(row + clampi(k - 2 + i, 0, bw - 1));
    }
    float o[S];
#pragma unroll
    for (int u = 0; u < S; ++u) {
      const float rx = rs * ((float)(k * S + u) + 0.5f) - 0.5f;
      const float fx = floorf(rx);
      float cx[4];
      cubic_coeffs(rx - fx, cx);
      const int off = u < S / 2 ? 0 : 1;       // floor(rx) = k - 1 for the first half of the S outputs, k for the second
      float val = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float rsum = v[j][off] * cx[0] + v[j][off + 1] * cx[1] + v[j][off + 2] * cx[2] + v[j][off + 3] * cx[3];
        val += rsum * cy[j];
      }
      o[u] = val;
    }
    float* dst = out + (size_t)oy * ow + (size_t)kx * S;
    if (S == 4) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    else *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[S - 1]);
  }
}

}  // namespace lfsr

using namespace lfsr;

extern "C" int lfsr_interp(const float* in, float* out, int n, int h, int w, int scale, int mode, int block_h,
                           int block_w, void* stream) {
  LFSR_REQUIRE(in && out, "lfsr_interp: null pointer");
  LFSR_REQUIRE(n > 0 && h > 0 && w > 0 && scale >= 1, "lfsr_interp: bad geometry");
  LFSR_REQUIRE(mode == LFSR_INTERP_BICUBIC || mode == LFSR_INTERP_BILINEAR, "lfsr_interp: bad mode %d", mode);
  LFSR_REQUIRE(block_h > 0 && block_w > 0 && h % block_h == 0 && w % block_w == 0,
               "lfsr_interp: block %dx%d does not tile %dx%d", block_h, block_w, h, w);
  LFSR_REQUIRE(n <= 65535 && (long long)h * scale * w * scale < 0x7fffffffLL, "lfsr_interp: image too large");
  if (mode == LFSR_INTERP_BICUBIC && (scale == 2 || scale == 4) && ((uintptr_t)out & 15) == 0) {
    const int per_t = h * scale * w;
    dim3 g2(ceil_div(per_t, 256) < 1184 ? ceil_div(per_t, 256) : 1184, n);
    if (scale == 4) interp_bicubic_kernel<4><<<g2, 256, 0, (cudaStream_t)stream>>>(in, out, h, w, block_h, block_w);
    else interp_bicubic_kernel<2><<<g2, 256, 0, (cudaStream_t)stream>>>(in, out, h, w, block_h, block_w);
    return check_launch("interp_bicubic_kernel");
  }
  const int per = h * scale * w * scale;
  dim3 grid(ceil_div(per, 256) < 1184 ? ceil_div(per, 256) : 1184, n);
  interp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, n, h, w, scale, mode, block_h, block_w);
  return check_launch("interp_kernel");
}


// ---- MATLAB-style imresize (utils/imresize.py): one separable pass in fp64 ------------------------------------------
// out[a][i][b] = sum_p w[i][p] * in[a][idx[i][p]][b] over a tensor viewed as [outer][length][inner]; the weights and the
// border-reflected indices of the resampled dimension come from the host (imresize.py:32-55, a few KB). Terms are added
// in tap order p = 0..P-1 with separate multiply and add (no FMA contraction), like the reference's numpy sum.
namespace lfsr {
__global__ void __launch_bounds__(256)
resample_f64_kernel(const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ w,
                    const int* __restrict__ idx, int in_len, int out_len, int inner, int P, long long total) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(t % inner);
    const long long r = t / inner;
    const int i = (int)(r % out_len);
    const long long a = r / out_len;
    const double* src = in + a * in_len * (long long)inner + b;
    double acc = 0.0;
    for (int p = 0; p < P; ++p) acc = __dadd_rn(acc, __dmul_rn(w[i * P + p], src[(long long)idx[i * P + p] * inner]));
    out[t] = acc;
  }
}
}  // namespace lfsr

extern "C" int lfsr_resample_f64(const double* in, double* out, const double* weights, const int32_t* indices, int outer,
                                 int in_len, int out_len, int inner, int taps, void* stream) {
  LFSR_REQUIRE(in && out && weights && indices, "lfsr_resample_f64: null pointer");
  LFSR_REQUIRE(outer > 0 && in_len > 0 && out_len > 0 && inner > 0 && taps > 0, "lfsr_resample_f64: bad geometry");
  const long long total = (long long)outer * out_len * inner;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  lfsr::resample_f64_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, weights, indices, in_len, out_len, inner, taps, total);
  return lfsr::check_launch("resample_f64_kernel");
}
