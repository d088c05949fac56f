// LFdivide / LFintegrate gather kernels (reference: utils/utils.py:137-178, train.py:300-319).
// Pure index permutations -> bit-exact by construction. HBM-bound: every output float is written
// once with 128-bit stores; source reads are contiguous (or mirrored-contiguous) 32-float runs.
#include <stdarg.h>
#include "lfsr_common.cuh"

namespace lfsr {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// symmetric (edge-repeating) mirror used by ImageExtend (utils/utils.py:137-149)
__device__ __forceinline__ int mirror(int i, int n) { return i < 0 ? -i - 1 : (i >= n ? 2 * n - 1 - i : i); }

// one thread = 4 consecutive x of one patch row (same view because P % 4 == 0);
// blockIdx.y walks the patches of the shard so all index math is 32-bit
__global__ void __launch_bounds__(256)
divide_kernel(const float* __restrict__ scene, float* __restrict__ patches, int A, int h0, int w0, int P,
              int S, int bdr, int numV, int u_begin, int npatch) {
  const int row = A * P;            // floats per patch row
  const int row4 = row >> 2;
  const int per_patch4 = row * row4;
  const int sw = A * w0;            // scene row stride
  for (int pidx = blockIdx.y; pidx < npatch; pidx += gridDim.y) {
    const int n2 = pidx % numV;
    const int n1 = pidx / numV + u_begin;
    float4* dst = reinterpret_cast<float4*>(patches) + (size_t)pidx * per_patch4;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per_patch4; t += gridDim.x * blockDim.x) {
      const int x4 = t % row4;
      const int py = t / row4;
      const int a1 = py / P, y = py - a1 * P;
      const int px = x4 << 2;
      const int a2 = px / P, x = px - a2 * P;
      const int sy = mirror(n1 * S + y - bdr, h0);
      const float* src = scene + (size_t)(a1 * h0 + sy) * sw + a2 * w0;
      const int gx = n2 * S + x - bdr;
      float4 v;
      v.x = __ldg(src + mirror(gx, w0));
      v.y = __ldg(src + mirror(gx + 1, w0));
      v.z = __ldg(src + mirror(gx + 2, w0));
      v.w = __ldg(src + mirror(gx + 3, w0));
      dst[t] = v;
    }
  }
}

// blockIdx.y walks (view row a1, output row Y); x covers (a2, X / VEC)
template <int VEC>
__global__ void __launch_bounds__(256)
integrate_kernel(const float* __restrict__ patches, float* __restrict__ out, int A, int pz, int ss, int h,
                 int w, int numV, int u_begin, int y_begin, int y_count) {
  const int wv = w / VEC;           // vectors per view row
  const int bdr = (pz - ss) / 2;
  const size_t prow = (size_t)A * pz;
  for (int ry = blockIdx.y; ry < A * y_count; ry += gridDim.y) {
    const int a1 = ry / y_count;
    const int Y = y_begin + (ry - a1 * y_count);
    const int n1 = Y / ss;
    const float* prow_base = patches + (size_t)(n1 - u_begin) * numV * prow * prow +
                             (size_t)(a1 * pz + bdr + (Y - n1 * ss)) * prow;
    float* orow = out + (size_t)(a1 * h + Y) * ((size_t)A * w);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < A * wv; t += gridDim.x * blockDim.x) {
      const int a2 = t / wv;
      const int X = (t - a2 * wv) * VEC;
      const int n2 = X / ss;
      const float* src = prow_base + (size_t)n2 * prow * prow + a2 * pz + bdr + (X - n2 * ss);
      float* dst = orow + (size_t)a2 * w + X;
      if (VEC == 4) *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
      else *dst = __ldg(src);
    }
  }
}


}  // namespace lfsr

using namespace lfsr;

extern "C" const char* lfsr_last_error(void) { return g_err; }
extern "C" int lfsr_abi_version(void) { return LFSR_ABI_VERSION; }
extern "C" int lfsr_built_for_sm100a(void) { return 1; }
extern "C" uint64_t lfsr_launch_count(void) { return g_launches.load(); }

extern "C" int lfsr_divide_rows(const float* scene, float* patches, int ang, int h0, int w0, int patch,
                                int stride, int u_begin, int u_end, void* stream) {
  LFSR_REQUIRE(scene && patches, "lfsr_divide: null pointer");
  LFSR_REQUIRE(ang > 0 && h0 > 0 && w0 > 0 && patch > 0 && stride > 0 && patch >= stride,
               "lfsr_divide: bad geometry A=%d h0=%d w0=%d P=%d S=%d", ang, h0, w0, patch, stride);
  LFSR_REQUIRE(patch % 4 == 0, "lfsr_divide: patch size %d must be a multiple of 4", patch);
  const int bdr = (patch - stride) / 2;
  const int numU = (h0 + 2 * bdr - 1) / stride, numV = (w0 + 2 * bdr - 1) / stride;
  // The reference pads bdr above and bdr + stride - 1 below out of ONE mirrored copy of the view (ImageExtend,
  // utils/utils.py:141-147: at most n rows per side) and its unfold must then yield exactly numU x numV windows
  // (utils.py:160-164), else it raises: same accept / reject rule here (oracle/make_golden.py sweeps it against the
  // reference). Accepted geometries never index past one reflection, which is all mirror() implements.
  auto tiles = [&](int n, int want) {
    const int below = bdr + stride - 1 < n ? bdr + stride - 1 : n;
    const int ext = n + bdr + below;
    const int windows = ext >= patch ? (ext - patch) / stride + 1 : 0;
    return n >= bdr && windows >= 1 && windows == want;
  };
  LFSR_REQUIRE(tiles(h0, numU) && tiles(w0, numV),
               "lfsr_divide: %dx%d views cannot be tiled with patch %d / stride %d (the reference's LFdivide raises too)", h0,
               w0, patch, stride);
  LFSR_REQUIRE(0 <= u_begin && u_begin <= u_end && u_end <= numU, "lfsr_divide: row shard [%d,%d) outside [0,%d)",
               u_begin, u_end, numU);
  if (u_begin == u_end) return LFSR_OK;
  const int npatch = (u_end - u_begin) * numV;
  const int per_patch4 = (ang * patch) * (ang * patch / 4);
  dim3 grid(ceil_div(per_patch4, 256 * 4), npatch < 32768 ? npatch : 32768);
  divide_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(scene, patches, ang, h0, w0, patch, stride, bdr, numV, u_begin,
                                                        npatch);
  return check_launch("divide_kernel");
}

extern "C" int lfsr_divide(const float* scene, float* patches, int ang, int h0, int w0, int patch, int stride,
                           void* stream) {
  if (patch < stride || stride <= 0) { set_error("lfsr_divide: bad patch/stride"); return LFSR_ERR_INVALID; }
  const int bdr = (patch - stride) / 2;
  const int numU = (h0 + 2 * bdr - 1) / stride;
  return lfsr_divide_rows(scene, patches, ang, h0, w0, patch, stride, 0, numU, stream);
}

extern "C" int lfsr_integrate_rows(const float* patches, float* out, int ang, int pz, int stride, int h, int w,
                                   int num_u, int num_v, int u_begin, int u_end, void* stream) {
  LFSR_REQUIRE(patches && out, "lfsr_integrate: null pointer");
  LFSR_REQUIRE(ang > 0 && pz >= stride && stride > 0 && h > 0 && w > 0, "lfsr_integrate: bad geometry");
  LFSR_REQUIRE(h <= num_u * stride && w <= num_v * stride,
               "lfsr_integrate: output %dx%d larger than the stitched grid %dx%d", h, w, num_u * stride,
               num_v * stride);
  LFSR_REQUIRE(0 <= u_begin && u_begin <= u_end && u_end <= num_u, "lfsr_integrate: bad row shard");
  int y_begin = u_begin * stride;
  int y_end = u_end * stride < h ? u_end * stride : h;
  if (y_end <= y_begin) return LFSR_OK;
  const int y_count = y_end - y_begin;
  const int bdr = (pz - stride) / 2;
  const bool vec = (w % 4 == 0) && (stride % 4 == 0) && (pz % 4 == 0) && (bdr % 4 == 0) &&
                   ((uintptr_t)patches % 16 == 0) && ((uintptr_t)out % 16 == 0);
  const int rows = ang * y_count;
  if (vec) {
    dim3 grid(ceil_div(ang * (w / 4), 256), rows < 32768 ? rows : 32768);
    integrate_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(patches, out, ang, pz, stride, h, w, num_v, u_begin,
                                                                y_begin, y_count);
  } else {
    dim3 grid(ceil_div(ang * w, 256), rows < 32768 ? rows : 32768);
    integrate_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(patches, out, ang, pz, stride, h, w, num_v, u_begin,
                                                                y_begin, y_count);
  }
  return check_launch("integrate_kernel");
}


// ---- colour tail of test() (train.py:329-341, utils/utils.py:191-204): Y mosaic + CbCr mosaic -> uint8 RGB views -------
// rgb8[u][v][y][x][c] = uint8(clip(M[c][0]*Y + M[c][1]*Cb + M[c][2]*Cr - off[c], 0, 1) * 255), fp64 with separate multiplies and
// adds in the reference's left-to-right order (no FMA contraction) and truncation like numpy's astype('uint8'); the
// (a1 h)(a2 w) -> a1 a2 h w view split is folded into the store address. 1 byte per channel leaves the GPU, not 4.
namespace lfsr {
struct ColourMat { double m[9]; double off[3]; };
__global__ void __launch_bounds__(256)
ycbcr_rgb8_kernel(const float* __restrict__ y, const float* __restrict__ cb, const float* __restrict__ cr,
                  unsigned char* __restrict__ rgb, int ang, int h, int w, ColourMat cm) {
  const int W = ang * w;
  const long long total = (long long)ang * h * W;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(t % W), Y = (int)(t / W);
    const int u = Y / h, yy = Y - u * h, v = X / w, xx = X - v * w;
    const double a = (double)y[t], b = (double)cb[t], c = (double)cr[t];
    unsigned char* dst = rgb + ((((long long)u * ang + v) * h + yy) * w + xx) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double r = __dadd_rn(__dadd_rn(__dmul_rn(cm.m[3 * k], a), __dmul_rn(cm.m[3 * k + 1], b)), __dmul_rn(cm.m[3 * k + 2], c));
      r = __dadd_rn(r, -cm.off[k]);
      r = r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);          // numpy clip (NaN propagates there; inputs here are finite)
      dst[k] = (unsigned char)(int)__dmul_rn(r, 255.0);
    }
  }
}
}  // namespace lfsr

extern "C" int lfsr_ycbcr_to_rgb8(const float* y, const float* cb, const float* cr, unsigned char* rgb, int ang, int h, int w,
                                  const double* mat_inv255, const double* offset, void* stream) {
  LFSR_REQUIRE(y && cb && cr && rgb && mat_inv255 && offset, "lfsr_ycbcr_to_rgb8: null pointer");
  LFSR_REQUIRE(ang > 0 && h > 0 && w > 0, "lfsr_ycbcr_to_rgb8: bad geometry");
  lfsr::ColourMat cm;
  for (int i = 0; i < 9; ++i) cm.m[i] = mat_inv255[i];
  for (int i = 0; i < 3; ++i) cm.off[i] = offset[i];
  const long long total = (long long)ang * h * ang * w;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  lfsr::ycbcr_rgb8_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(y, cb, cr, rgb, ang, h, w, cm);
  return lfsr::check_launch("ycbcr_rgb8_kernel");
}
