// fp32 CUDA-core convolution family: generic implicit-GEMM conv with the fused epilogue of
// lfsr_conv_desc, and depthwise conv. These carry every layer the tcgen05 path
// (lfsr_conv_tc.cu) does not take: 1-channel stems/heads, the 18/16/13-channel branches of
// MyEfficientLFNet.py:119-327, strided A x A "angular" and 1 x A*A "EPI" convs
// (DistgSSR.py:85-97, LF_InterNet.py:48-55).
//
// Implicit GEMM: M = output pixels (2-D tile), N = output channels, K = (tap, cin) flattened.
// 256 threads, BK = 8, register-prefetched double-buffered smem, TMxTN register tiles.
#include "lfsr_common.cuh"

namespace lfsr {

struct ConvArgs {
  TView in, out, mul, res;
  const float* w;
  const float* bias;
  const float* in_scale;
  long long in_scale_ld;
  int kh, kw, sh, sw, dh, dw, ph, pw;
  int in_perm, out_perm, A;
  int ry, rx, shuf_mode;
  int bh, bw;  // view blocking (0 = off)
  int act, mul_act;
  float slope, alpha;
  int OH, OW, cin, cout, K;
  int tiles_x, tiles_y;
};

constexpr int BK = 8;

template <int TILE_H, int TILE_W, int BN, int TM, int TN, bool VEC>
__global__ void __launch_bounds__(256)
conv_igemm_f32(const ConvArgs a) {
  constexpr int BM = TILE_H * TILE_W;
  constexpr int LDA = BM + 4;
  constexpr int NTN = BN / TN;          // thread columns
  constexpr int GROUPS = 256 / BM;      // loader thread groups per pixel (1 or 2)
  constexpr int KPT = BK / GROUPS;      // k elements per loader thread
  static_assert((BM / TM) * NTN == 256, "thread tiling must cover the block tile");
  static_assert(BM == 128 || BM == 256, "BM");
  static_assert(KPT % 4 == 0, "loader granularity");
  constexpr int WPT = (BK * BN + 255) / 256;  // weight elements per thread

  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][BN];

  const int tid = threadIdx.x;
  int tile = blockIdx.x;
  const int tx0 = (tile % a.tiles_x) * TILE_W; tile /= a.tiles_x;
  const int ty0 = (tile % a.tiles_y) * TILE_H;
  const int img = tile / a.tiles_y;
  const int n0 = blockIdx.y * BN;

  // ---- loader role: one pixel, KPT consecutive k per chunk
  const int lp = tid % BM;
  const int lg = tid / BM;
  const int loy = ty0 + lp / TILE_W, lox = tx0 + lp % TILE_W;
  const bool lvalid = loy < a.OH && lox < a.OW;
  const int iy0 = loy * a.sh - a.ph, ix0 = lox * a.sw - a.pw;
  const float* scale_row = a.in_scale ? a.in_scale + (size_t)img * a.in_scale_ld : nullptr;

  float areg[KPT];
  float breg[WPT];
  const int lby = a.bh > 0 ? loy / a.bh : 0, lbx = a.bw > 0 ? lox / a.bw : 0;
  auto in_block = [&](int iy, int ix) {
    return (a.bh <= 0 || iy / a.bh == lby) && (a.bw <= 0 || ix / a.bw == lbx);
  };

  auto load_a = [&](int kc) {
    const int kbase = kc * BK + lg * KPT;
#pragma unroll
    for (int q = 0; q < KPT; q += 4) {
      const int k = kbase + q;
      float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
      if (VEC) {
        if (lvalid && k < a.K) {
          int tap = k / a.cin, ci = k - tap * a.cin;
          int ky = tap / a.kw, kx = tap - ky * a.kw;
          int iy = iy0 + ky * a.dh, ix = ix0 + kx * a.dw;
          if (iy >= 0 && iy < a.in.h && ix >= 0 && ix < a.in.w && in_block(iy, ix)) {
            if (a.in_perm) { int sy, sx; macpi_to_sai(iy, ix, a.A, a.in.h, a.in.w, sy, sx); iy = sy; ix = sx; }
            const float4 t = __ldg(reinterpret_cast<const float4*>(a.in.p + a.in.pix(img, iy, ix) + ci));
            v0 = t.x; v1 = t.y; v2 = t.z; v3 = t.w;
            if (scale_row) {
              v0 *= __ldg(scale_row + ci); v1 *= __ldg(scale_row + ci + 1);
              v2 *= __ldg(scale_row + ci + 2); v3 *= __ldg(scale_row + ci + 3);
            }
          }
        }
      } else {
        float vv[4] = {0.f, 0.f, 0.f, 0.f};
        if (lvalid) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int ke = k + e;
            if (ke < a.K) {
              int tap = ke / a.cin, ci = ke - tap * a.cin;
              int ky = tap / a.kw, kx = tap - ky * a.kw;
              int iy = iy0 + ky * a.dh, ix = ix0 + kx * a.dw;
              if (iy >= 0 && iy < a.in.h && ix >= 0 && ix < a.in.w && in_block(iy, ix)) {
                if (a.in_perm) { int sy, sx; macpi_to_sai(iy, ix, a.A, a.in.h, a.in.w, sy, sx); iy = sy; ix = sx; }
                float t = __ldg(a.in.p + a.in.pix(img, iy, ix) + ci);
                if (scale_row) t *= __ldg(scale_row + ci);
                vv[e] = t;
              }
            }
          }
        }
        v0 = vv[0]; v1 = vv[1]; v2 = vv[2]; v3 = vv[3];
      }
      areg[q] = v0; areg[q + 1] = v1; areg[q + 2] = v2; areg[q + 3] = v3;
    }
  };
  auto load_b = [&](int kc) {
#pragma unroll
    for (int r = 0; r < WPT; ++r) {
      const int idx = tid + r * 256;
      float v = 0.f;
      if (idx < BK * BN) {
        const int kk = idx / BN, nn = idx - kk * BN;
        const int k = kc * BK + kk, co = n0 + nn;
        if (k < a.K && co < a.cout) v = __ldg(a.w + (size_t)k * a.cout + co);
      }
      breg[r] = v;
    }
  };
  auto store_ab = [&](int buf) {
#pragma unroll
    for (int q = 0; q < KPT; ++q) As[buf][lg * KPT + q][lp] = areg[q];
#pragma unroll
    for (int r = 0; r < WPT; ++r) {
      const int idx = tid + r * 256;
      if (idx < BK * BN) Bs[buf][idx / BN][idx % BN] = breg[r];
    }
  };

  // ---- compute role
  const int tn = tid % NTN, tm = tid / NTN;
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = (a.K + BK - 1) / BK;
  load_a(0); load_b(0);
  store_ab(0);
  __syncthreads();
  for (int kc = 0; kc < nk; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nk) { load_a(kc + 1); load_b(kc + 1); }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], bv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][tm * TM + i]);
        av[i] = t.x; av[i + 1] = t.y; av[i + 2] = t.z; av[i + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][tn * TN + j]);
        bv[j] = t.x; bv[j + 1] = t.y; bv[j + 2] = t.z; bv[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kc + 1 < nk) store_ab(buf ^ 1);
    __syncthreads();
  }

  // ---- fused epilogue
  const int r2 = a.ry * a.rx;
  const int cq = a.cout / r2;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = tm * TM + i;
    const int oy = ty0 + m / TILE_W, ox = tx0 + m % TILE_W;
    if (oy >= a.OH || ox >= a.OW) continue;
    int py = oy, px = ox;
    if (a.out_perm) macpi_to_sai(oy, ox, a.A, a.OH, a.OW, py, px);
    const size_t mul_base = a.mul.p ? a.mul.pix(img, oy, ox) : 0;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tn * TN + j;
      if (co >= a.cout) continue;
      float v = acc[i][j];
      if (a.bias) v += __ldg(a.bias + co);
      v = apply_act(v, a.act, a.slope);
      if (a.mul.p) v *= apply_act(__ldg(a.mul.p + mul_base + co), a.mul_act, 0.f);
      v *= a.alpha;
      int sy = py, sx = px, sc = co;
      if (r2 > 1) {
        int sub;
        if (a.shuf_mode == LFSR_SHUF_CHANNEL_MAJOR) { sc = co / r2; sub = co - sc * r2; }
        else { sub = co / cq; sc = co - sub * cq; }
        const int si = sub / a.rx, sj = sub - si * a.rx;
        sy = py * a.ry + si; sx = px * a.rx + sj;
      }
      const size_t o = a.out.pix(img, sy, sx) + sc;
      if (a.res.p) v += a.res.p[a.res.pix(img, sy, sx) + sc];
      a.out.p[o] = v;
    }
  }
}

template <int TILE_H, int TILE_W, int BN, int TM, int TN>
static int launch_conv(const ConvArgs& a, bool vec, cudaStream_t st) {
  dim3 grid(a.in.n * a.tiles_x * a.tiles_y, ceil_div(a.cout, BN));
  if (vec) conv_igemm_f32<TILE_H, TILE_W, BN, TM, TN, true><<<grid, 256, 0, st>>>(a);
  else conv_igemm_f32<TILE_H, TILE_W, BN, TM, TN, false><<<grid, 256, 0, st>>>(a);
  return check_launch("conv_igemm_f32");
}

}  // namespace lfsr

using namespace lfsr;

extern "C" int lfsr_conv2d_f32(const lfsr_tensor* in, const float* w_packed, const lfsr_tensor* out,
                               const lfsr_conv_desc* d, void* stream) {
  LFSR_REQUIRE(tensor_ok(in) && tensor_ok(out) && w_packed && d, "lfsr_conv2d_f32: null/invalid tensor");
  LFSR_REQUIRE(d->kh > 0 && d->kw > 0 && d->stride_h > 0 && d->stride_w > 0 && d->dil_h > 0 && d->dil_w > 0 &&
                   d->pad_h >= 0 && d->pad_w >= 0,
               "lfsr_conv2d_f32: bad geometry");
  const int ry = d->shuf_ry > 0 ? d->shuf_ry : 1, rx = d->shuf_rx > 0 ? d->shuf_rx : 1;
  const int OH = (in->h + 2 * d->pad_h - d->dil_h * (d->kh - 1) - 1) / d->stride_h + 1;
  const int OW = (in->w + 2 * d->pad_w - d->dil_w * (d->kw - 1) - 1) / d->stride_w + 1;
  LFSR_REQUIRE(OH > 0 && OW > 0, "lfsr_conv2d_f32: empty output");
  LFSR_REQUIRE(out->n == in->n && out->h == OH * ry && out->w == OW * rx,
               "lfsr_conv2d_f32: out %dx%dx%d does not match conv %dx%d shuffle %dx%d", out->n, out->h, out->w, OH, OW,
               ry, rx);
  const int cout = out->c * ry * rx;
  LFSR_REQUIRE((d->in_perm == 0 && d->out_perm == 0) || d->perm_a > 0, "lfsr_conv2d_f32: perm needs perm_a");
  if (d->block_h > 0 || d->block_w > 0)
    LFSR_REQUIRE(d->stride_h == 1 && d->stride_w == 1 && d->in_perm == 0 && OH == in->h && OW == in->w,
                 "lfsr_conv2d_f32: view blocking needs a stride-1 'same' convolution");
  if (d->in_perm) LFSR_REQUIRE(in->h % d->perm_a == 0 && in->w % d->perm_a == 0, "lfsr_conv2d_f32: in_perm geometry");
  if (d->out_perm) LFSR_REQUIRE(OH % d->perm_a == 0 && OW % d->perm_a == 0, "lfsr_conv2d_f32: out_perm geometry");
  if (d->mul.ptr)
    LFSR_REQUIRE(d->mul.n == in->n && d->mul.h == OH && d->mul.w == OW && d->mul.c == cout && d->mul.ld >= cout,
                 "lfsr_conv2d_f32: mul tensor geometry");
  if (d->res.ptr)
    LFSR_REQUIRE(d->res.n == out->n && d->res.h == out->h && d->res.w == out->w && d->res.c == out->c &&
                     d->res.ld >= out->c,
                 "lfsr_conv2d_f32: res tensor geometry");

  ConvArgs a;
  a.in = view_of(in); a.out = view_of(out);
  a.mul = d->mul.ptr ? view_of(&d->mul) : null_view();
  a.res = d->res.ptr ? view_of(&d->res) : null_view();
  a.w = w_packed; a.bias = d->bias; a.in_scale = d->in_scale;
  a.in_scale_ld = d->in_scale_ld > 0 ? d->in_scale_ld : in->c;
  a.kh = d->kh; a.kw = d->kw; a.sh = d->stride_h; a.sw = d->stride_w; a.dh = d->dil_h; a.dw = d->dil_w;
  a.ph = d->pad_h; a.pw = d->pad_w;
  a.in_perm = d->in_perm; a.out_perm = d->out_perm; a.A = d->perm_a > 0 ? d->perm_a : 1;
  a.ry = ry; a.rx = rx; a.shuf_mode = d->shuf_mode;
  a.bh = d->block_h; a.bw = d->block_w;
  a.act = d->act; a.slope = d->act_slope; a.alpha = d->alpha; a.mul_act = d->mul_act;
  a.OH = OH; a.OW = OW; a.cin = in->c; a.cout = cout; a.K = d->kh * d->kw * in->c;
  const bool vec = (in->c % 4 == 0) && (in->ld % 4 == 0) && ((uintptr_t)in->ptr % 16 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (cout > 32) {
    a.tiles_x = ceil_div(OW, 16); a.tiles_y = ceil_div(OH, 8);
    return launch_conv<8, 16, 64, 8, 4>(a, vec, st);
  } else if (cout > 16) {
    a.tiles_x = ceil_div(OW, 16); a.tiles_y = ceil_div(OH, 16);
    return launch_conv<16, 16, 32, 8, 4>(a, vec, st);
  } else {
    a.tiles_x = ceil_div(OW, 16); a.tiles_y = ceil_div(OH, 16);
    return launch_conv<16, 16, 16, 4, 4>(a, vec, st);
  }
}
