// Depthwise convolutions (groups == channels) of the SR networks, NHWC fp32, sm_100a.
//
//   dw_tile_kernel   channel counts that are multiples of 4 on 16-byte aligned rows (the 64- and 60-channel trunks):
//                    one CTA = one 16x32 output tile x (up to) 16 channels of one branch. The tile plus its dilation
//                    halo is staged once in shared memory with 16-byte cp.async (out-of-image pixels zero
//                    filled = the conv's zero padding), each thread then owns one channel quad (float4) of one
//                    pixel column and walks the tile rows; a warp reads 8 pixels x 64 B = 512 contiguous bytes
//                    per tap, conflict-free. Several branches (different taps / dilations / channel windows)
//                    share one launch: MultiScaleSpatial's 1/3/5/7 kernels on four 16-channel slices and
//                    FastConvSSM's four dilations of one input (MyEfficientLFNetV4_5.py:218-221, :268-271). With
//                    DwParams::sa the epilogue is the SA-modulator tail of the Track-2 model (lfsr_sa_modulate).
//   dwconv_kernel    anything else (18-channel groups of the Track-2 model): one thread per pixel x channel.
//
// HBM-bound: algorithmic traffic is one read of the input window and one write per branch output.
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include "lfsr_common.cuh"

namespace lfsr {

__global__ void __launch_bounds__(256)
dwconv_kernel(TView in, TView out, const float* __restrict__ w, const float* __restrict__ scale,
              const float* __restrict__ shift, int kh, int kw, int dh, int dw, int act, float slope) {
  const int C = in.c;
  const int img = blockIdx.y;                 // 32-bit index math inside one image
  const int per = in.h * in.w * C;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < per; t += gridDim.x * blockDim.x) {
    const int c = t % C;
    const int r = t / C;
    const int x = r % in.w;
    const int y = r / in.w;
    const int ph = (kh / 2) * dh, pw = (kw / 2) * dw;
    float acc = 0.f;
    for (int ky = 0; ky < kh; ++ky) {
      const int iy = y - ph + ky * dh;
      if (iy < 0 || iy >= in.h) continue;
      for (int kx = 0; kx < kw; ++kx) {
        const int ix = x - pw + kx * dw;
        if (ix < 0 || ix >= in.w) continue;
        acc = fmaf(__ldg(in.p + in.pix(img, iy, ix) + c), __ldg(w + (ky * kw + kx) * C + c), acc);
      }
    }
    if (scale) acc = acc * __ldg(scale + c) + __ldg(shift + c);
    out.p[out.pix(img, y, x) + c] = apply_act(acc, act, slope);
  }
}

constexpr int DW_TH = 16, DW_TW = 32, DW_CH = 16, DW_MAXB = 8;

struct DwBranch {
  const float* w;
  const float* scale;
  const float* shift;
  int kh, kw, dh, dw;
  int in_c0, out_c0, c;
  int act;
  float slope;
  int item0;              // first grid.y index of this branch (one item = one 16-channel chunk)
};

struct DwParams {
  TView in, out;
  DwBranch br[DW_MAXB];
  int nbr, tiles_x, w_floats;   // w_floats: shared-memory floats reserved for the taps ahead of the tile
  // SA-modulator tail (MyEfficientLFNet.py:495-515, :207): with s = act(affine(dw(x))) the kernel stores
  // x * (sa_w0 * s + sa_w1 * amod[view of the pixel]) + res instead of s
  int sa;
  float sa_w0, sa_w1;
  __half* o16;                 // SA tail: optional fp16 copy of output channels [0, o16_c) (operand of the next tensor-core layer)
  int o16_ld, o16_c;
  int o16_skip_lo, o16_skip_hi;   // channels in [lo, hi) are left out of the fp16 copy (nobody reads them)
  TView amod, res;
  // use_tma: the halo'd tile of branch b is fetched by ONE thread with a 4-D tensor load through tm[b] (box = 16 channels
  // x tile+halo, out-of-image elements zero filled) instead of ~25 cp.async per thread: staging cost as many
  // instructions as the stencil itself and the kernel is issue-bound at its low (shared-memory limited) occupancy
  int use_tma, tile_floats;
  int res_tma;                 // sa_tile_kernel: the residual tile is fetched through tm[1] next to the input tile
  // share: all branches read the same channel window (FastConvSSM's four dilations of one tensor): a CTA stages ONE tile
  // with the largest halo (branch share_b) and runs every branch on it; grid.y then enumerates channel chunks only
  int share, share_b;
  CUtensorMap tm[DW_MAXB];
};

__device__ __forceinline__ void cp_async16_zfill(float* dst, const float* src, bool ok) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  const int sz = ok ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}

__device__ __forceinline__ uint32_t dw_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dw_mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = dw_smem_u32(bar);
  uint32_t done = 0;
  for (uint32_t it = 0; it < (1u << 26); ++it) {      // bounded: a protocol bug must fault, not hang the GPU
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}

// float4 accumulators live as two packed pairs: acc += v * w costs two FFMA2
struct Acc4 { f32x2 lo, hi; };
__device__ __forceinline__ void fma4(Acc4& a, const float4& v, const float4& w) {
  a.lo = fma2(pack2(v.x, v.y), pack2(w.x, w.y), a.lo);
  a.hi = fma2(pack2(v.z, v.w), pack2(w.z, w.w), a.hi);
}

template <int KH, int KW>   // 0,0 = runtime kernel size, taps read from shared memory
__device__ __forceinline__ void dw_compute(const DwParams& p, const DwBranch& B, const float* wS, const float* tS,
                                           int SW, int img, int ty0, int tx0, int cout0, int wofs) {
  const int tid = threadIdx.x;
  const int q = tid & 3, lx = (tid >> 2) & 31, ly0 = tid >> 7;
  const int kh = KH ? KH : B.kh, kw = KW ? KW : B.kw;
  const int ox = tx0 + lx;
  if (ox >= p.in.w || wofs + q * 4 >= B.c) return;
  float4 wr[KH * KW ? KH * KW : 1];
  if (KH) {
#pragma unroll
    for (int t = 0; t < KH * KW; ++t) wr[t] = *reinterpret_cast<const float4*>(wS + t * DW_CH + q * 4);
  }
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (B.scale) {
    sc = __ldg(reinterpret_cast<const float4*>(B.scale + wofs + q * 4));
    sh = __ldg(reinterpret_cast<const float4*>(B.shift + wofs + q * 4));
  }
  const int row_f = SW * DW_CH;
  const int dyf = B.dh * row_f, dxf = B.dw * DW_CH;
  // SA tail: view column of this thread's pixel column, and the view height as a float (row / vh for rows < 2^20 is
  // exact through (row + 0.5) / vh in fp32)
  const int sa_ax = p.sa ? ox / (p.in.w / p.amod.w) : 0;
  const float sa_vh = p.sa ? (float)(p.in.h / p.amod.h) : 1.f;
  for (int r = ly0; r < DW_TH; r += 2) {
    const int oy = ty0 + r;
    if (oy >= p.in.h) break;
    const float* base = tS + r * row_f + lx * DW_CH + q * 4;
    Acc4 a2;
    a2.lo = pack2(0.f, 0.f); a2.hi = a2.lo;
    if (KH) {
#pragma unroll
      for (int ky = 0; ky < KH; ++ky)
#pragma unroll
        for (int kx = 0; kx < KW; ++kx)
          fma4(a2, *reinterpret_cast<const float4*>(base + ky * dyf + kx * dxf), wr[ky * KW + kx]);
    } else {
      for (int ky = 0; ky < kh; ++ky) {
        const float* rowp = base + ky * dyf;
        const float* wp = wS + ky * kw * DW_CH + q * 4;
#pragma unroll 7
        for (int kx = 0; kx < kw; ++kx)
          fma4(a2, *reinterpret_cast<const float4*>(rowp + kx * dxf), *reinterpret_cast<const float4*>(wp + kx * DW_CH));
      }
    }
    float4 acc;
    unpack2(a2.lo, acc.x, acc.y);
    unpack2(a2.hi, acc.z, acc.w);
    if (B.scale) {
      acc.x = acc.x * sc.x + sh.x; acc.y = acc.y * sc.y + sh.y; acc.z = acc.z * sc.z + sh.z; acc.w = acc.w * sc.w + sh.w;
    }
    if (B.act) {
      acc.x = apply_act(acc.x, B.act, B.slope); acc.y = apply_act(acc.y, B.act, B.slope);
      acc.z = apply_act(acc.z, B.act, B.slope); acc.w = apply_act(acc.w, B.act, B.slope);
    }
    if (p.sa) {
      const float4 xc = *reinterpret_cast<const float4*>(base + (kh / 2) * dyf + (kw / 2) * dxf);
      const float4 am = __ldg(reinterpret_cast<const float4*>(p.amod.p + p.amod.pix(img, (int)__fdividef((float)oy + 0.5f, sa_vh), sa_ax) +
                                                              cout0 + q * 4));
      acc.x = xc.x * (p.sa_w0 * acc.x + p.sa_w1 * am.x); acc.y = xc.y * (p.sa_w0 * acc.y + p.sa_w1 * am.y);
      acc.z = xc.z * (p.sa_w0 * acc.z + p.sa_w1 * am.z); acc.w = xc.w * (p.sa_w0 * acc.w + p.sa_w1 * am.w);
      if (p.res.p) {
        const float4 r = *reinterpret_cast<const float4*>(p.res.p + p.res.pix(img, oy, ox) + cout0 + q * 4);
        acc.x += r.x; acc.y += r.y; acc.z += r.z; acc.w += r.w;
      }
    }
    *reinterpret_cast<float4*>(p.out.p + p.out.pix(img, oy, ox) + cout0 + q * 4) = acc;
    if (p.o16 && cout0 + q * 4 < p.o16_c && !(cout0 + q * 4 >= p.o16_skip_lo && cout0 + q * 4 < p.o16_skip_hi)) {
      const __half2 h0 = __floats2half2_rn(acc.x, acc.y), h1 = __floats2half2_rn(acc.z, acc.w);
      uint2 v;
      v.x = *reinterpret_cast<const uint32_t*>(&h0); v.y = *reinterpret_cast<const uint32_t*>(&h1);
      *reinterpret_cast<uint2*>(p.o16 + ((size_t)((size_t)img * p.out.h + oy) * p.out.w + ox) * (size_t)p.o16_ld + cout0 + q * 4) = v;
    }
  }
}

// Large undilated kernels (5x5, 7x7 of MultiScaleSpatial): a thread owns 8 CONSECUTIVE output rows of one pixel column and
// channel quad. Per kernel column kx the K row-weights sit in registers and each of the 8 + K - 1 input rows is read once
// and accumulated into every output row it contributes to: 18 shared-memory reads per output at K = 7 instead of 98
// (dilated 3x3: 4.5 / 6 instead of 9 at dilation 2 / 4; nothing to gain at dilation 8).
template <int K, int D>      // K x K taps, dilation D (compile time: the row bookkeeping must unroll)
__device__ __forceinline__ void dw_compute_col(const DwParams& p, const DwBranch& B, const float* wS, const float* tS,
                                               int SW, int img, int ty0, int tx0, int cout0, int wofs) {
  const int tid = threadIdx.x;
  const int q = tid & 3, lx = (tid >> 2) & 31, r0 = (tid >> 7) * 8;
  const int ox = tx0 + lx;
  if (ox >= p.in.w || wofs + q * 4 >= B.c) return;
  const int row_f = SW * DW_CH;
  Acc4 acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) { acc[r].lo = pack2(0.f, 0.f); acc[r].hi = acc[r].lo; }
  const float* col = tS + r0 * row_f + lx * DW_CH + q * 4;
#pragma unroll 1
  for (int kx = 0; kx < K; ++kx) {
    float4 w[K];
#pragma unroll
    for (int ky = 0; ky < K; ++ky) w[ky] = *reinterpret_cast<const float4*>(wS + (ky * K + kx) * DW_CH + q * 4);
    const float* cp = col + kx * D * DW_CH;
#pragma unroll
    for (int ir = 0; ir < 8 + (K - 1) * D; ++ir) {
      const float4 v = *reinterpret_cast<const float4*>(cp + ir * row_f);
#pragma unroll
      for (int ky = 0; ky < K; ++ky) {
        const int r = ir - ky * D;                  // output row (relative to r0) this input row feeds through tap ky
        if (r >= 0 && r < 8) fma4(acc[r], v, w[ky]);
      }
    }
  }
  float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
  if (B.scale) {
    sc = __ldg(reinterpret_cast<const float4*>(B.scale + wofs + q * 4));
    sh = __ldg(reinterpret_cast<const float4*>(B.shift + wofs + q * 4));
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int oy = ty0 + r0 + r;
    if (oy >= p.in.h) break;
    float4 o;
    unpack2(acc[r].lo, o.x, o.y);
    unpack2(acc[r].hi, o.z, o.w);
    if (B.scale) { o.x = o.x * sc.x + sh.x; o.y = o.y * sc.y + sh.y; o.z = o.z * sc.z + sh.z; o.w = o.w * sc.w + sh.w; }
    if (B.act) {
      o.x = apply_act(o.x, B.act, B.slope); o.y = apply_act(o.y, B.act, B.slope);
      o.z = apply_act(o.z, B.act, B.slope); o.w = apply_act(o.w, B.act, B.slope);
    }
    *reinterpret_cast<float4*>(p.out.p + p.out.pix(img, oy, ox) + cout0 + q * 4) = o;
  }
}

// tap weights of this CTA's 16-channel chunk into shared memory: [tap][16] per branch (all branches in share mode)
__device__ __forceinline__ void dw_stage_taps(const DwParams& p, const DwBranch& B, float* wS, int wofs, int tid) {
  int wo = 0;
  const int nrun = p.share ? p.nbr : 1;
  for (int j = 0; j < nrun; ++j) {
    const DwBranch& Bj = p.share ? p.br[j] : B;
    const int nt = Bj.kh * Bj.kw * DW_CH;
    for (int i = tid; i < nt; i += 256) wS[wo + i] = wofs + (i & 15) < Bj.c ? __ldg(Bj.w + (i >> 4) * Bj.c + wofs + (i & 15)) : 0.f;
    wo += nt;
  }
}

// SHARE = false: 3 CTAs per SM (the stencil is issue/latency bound, occupancy matters more than registers);
// SHARE = true (one staged tile, all branches): shared memory allows 2 CTAs anyway, so it keeps its registers
template <bool SHARE>
__global__ void __launch_bounds__(256, SHARE ? 2 : 3)
dw_tile_kernel(const __grid_constant__ DwParams p) {
  extern __shared__ __align__(16) float dw_smem[];
  const int tid = threadIdx.x;
  int b = 0;
  if (SHARE) b = p.share_b;
  else while (b + 1 < p.nbr && (int)blockIdx.y >= p.br[b + 1].item0) ++b;
  const DwBranch& B = p.br[b];                       // the branch whose geometry defines the staged tile
  const int chunk = SHARE ? (int)blockIdx.y : (int)blockIdx.y - B.item0;
  const int wofs = chunk * DW_CH;
  const int cin0 = B.in_c0 + wofs, cout0 = B.out_c0 + wofs;
  const int hy = (B.kh / 2) * B.dh, hx = (B.kw / 2) * B.dw;
  const int SH = DW_TH + 2 * hy, SW = DW_TW + 2 * hx;
  const int img = blockIdx.z;
  const int tyi = blockIdx.x / p.tiles_x;
  const int ty0 = tyi * DW_TH, tx0 = (blockIdx.x - tyi * p.tiles_x) * DW_TW;
  // tile first (TMA destination: 128-byte aligned), taps behind it
  float* tS = dw_smem + ((128u - (dw_smem_u32(dw_smem) & 127u)) & 127u) / 4;
  float* wS = tS + p.tile_floats;
  __shared__ uint64_t bar;
  const int taps = B.kh * B.kw;
  const int H = p.in.h, W = p.in.w;
  if (p.use_tma) {
    if (tid == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dw_smem_u32(&bar)) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dw_smem_u32(&bar)), "r"(SH * SW * DW_CH * 4) : "memory");
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(dw_smem_u32(tS)), "l"(&p.tm[b]), "r"(dw_smem_u32(&bar)), "r"(cin0), "r"(tx0 - hx), "r"(ty0 - hy), "r"(img)
                   : "memory");
    }
    if (p.sa && p.res.p) {
      // SA tail: the residual rows this thread will add are requested into L2 now, while the tile is in flight. The row loop
      // cannot hoist its residual loads over the stores of the row before (they may alias), so without this every row
      // pays a full DRAM round trip in sequence: 8 rows x ~1.5 us was the whole lifetime of a CTA.
      const int q = tid & 3, lx = (tid >> 2) & 31;
      const int ox = tx0 + lx;
      if (ox < W && wofs + q * 4 < B.c && !(q & 1)) {          // one request per 32-byte sector
#pragma unroll
        for (int r = tid >> 7; r < DW_TH; r += 2) {
          const int oy = ty0 + r;
          if (oy < H) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res.p + p.res.pix(img, oy, ox) + cout0 + q * 4));
        }
      }
    }
    dw_stage_taps(p, B, wS, wofs, tid);
    __syncthreads();                   // taps visible; the barrier was initialised before anybody polls it
    dw_mbar_wait(&bar, 0);
  } else {
    dw_stage_taps(p, B, wS, wofs, tid);
    // rows of the halo'd tile that lie inside the image; everything else is zero padding
    const int quads_per_row = SW * 4;
    for (int sy = tid >> 7; sy < SH; sy += 2) {       // 128 threads per tile row
      const int gy = ty0 - hy + sy;
      const bool yok = gy >= 0 && gy < H;
      const float* grow = p.in.p + p.in.pix(img, yok ? gy : 0, 0) + cin0;
      float* srow = tS + sy * SW * DW_CH;
      for (int i = tid & 127; i < quads_per_row; i += 128) {
        const int sx = i >> 2, qq = i & 3;
        const int gx = tx0 - hx + sx;
        const bool ok = yok && gx >= 0 && gx < W && wofs + qq * 4 < B.c;     // (the last chunk of a branch may be partial)
        cp_async16_zfill(srow + i * 4, ok ? grow + (size_t)gx * p.in.ld + qq * 4 : p.in.p, ok);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  const int nrun = SHARE ? p.nbr : 1;
  int wo = 0;
  for (int j = 0; j < nrun; ++j) {
    const DwBranch& Bj = SHARE ? p.br[j] : B;
    // a branch with a smaller halo starts further inside the staged tile
    const int oy_ = hy - (Bj.kh / 2) * Bj.dh, ox_ = hx - (Bj.kw / 2) * Bj.dw;
    const float* tj = tS + (oy_ * SW + ox_) * DW_CH;
    const int co = Bj.out_c0 + wofs;
    if (SHARE && Bj.kh == 3 && Bj.kw == 3 && Bj.dh == Bj.dw && Bj.dh == 1) dw_compute_col<3, 1>(p, Bj, wS + wo, tj, SW, img, ty0, tx0, co, wofs);
    else if (SHARE && Bj.kh == 3 && Bj.kw == 3 && Bj.dh == Bj.dw && Bj.dh == 2) dw_compute_col<3, 2>(p, Bj, wS + wo, tj, SW, img, ty0, tx0, co, wofs);
    else if (SHARE && Bj.kh == 3 && Bj.kw == 3 && Bj.dh == Bj.dw && Bj.dh == 4) dw_compute_col<3, 4>(p, Bj, wS + wo, tj, SW, img, ty0, tx0, co, wofs);
    else if (Bj.kh == 3 && Bj.kw == 3) dw_compute<3, 3>(p, Bj, wS + wo, tj, SW, img, ty0, tx0, co, wofs);
    else if (Bj.kh == 1 && Bj.kw == 1) dw_compute<1, 1>(p, Bj, wS + wo, tj, SW, img, ty0, tx0, co, wofs);
    else if (!p.sa && Bj.kh == 5 && Bj.kw == 5 && Bj.dh == 1 && Bj.dw == 1) dw_compute_col<5, 1>(p, Bj, wS + wo, tj, SW, img, ty0, tx0, co, wofs);
    else if (!p.sa && Bj.kh == 7 && Bj.kw == 7 && Bj.dh == 1 && Bj.dw == 1) dw_compute_col<7, 1>(p, Bj, wS + wo, tj, SW, img, ty0, tx0, co, wofs);
    else dw_compute<0, 0>(p, Bj, wS + wo, tj, SW, img, ty0, tx0, co, wofs);
    wo += Bj.kh * Bj.kw * DW_CH;
  }
}

// SA-modulator tail (MyEfficientLFNet.py:495-515, :207) as its own lean kernel: out = x * (w0 * sigmoid(BN(dw3x3(x))) + w1 * amod[view])
// + res, optional fp16 copy of the first o16_c channels. Same tiling and TMA staging as dw_tile_kernel, but the row loop is written
// for exactly this case: the nine tap weights, the BN affine and w1 * amod live in registers, addresses advance by a
// constant per row (32-bit), the residual and modulation of the NEXT row are loaded before the current row is stored (and all
// residual rows are requested into L2 while the tile is still in flight), and the sigmoid is ex2 + rcp. The generic kernel
// spent ~300 issued instructions per output row and quad on this tail (switch on the activation, 64-bit pixel offsets per
// row, a spilled loop counter); this one ~80.
__device__ __forceinline__ float sa_sigmoid(float v) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}

constexpr int SA_THREADS = 512;
__global__ void __launch_bounds__(SA_THREADS, 2)
sa_tile_kernel(const __grid_constant__ DwParams p) {
  extern __shared__ __align__(16) float dw_smem[];
  const int tid = threadIdx.x;
  const DwBranch& B = p.br[0];
  const int wofs = (int)blockIdx.y * DW_CH;      // (channel chunk as the FASTEST grid index was tried: 0.332 vs 0.320 ms)
  const int cin0 = B.in_c0 + wofs, cout0 = B.out_c0 + wofs;
  const int d = B.dh;
  const int SH = DW_TH + 2 * d, SW = DW_TW + 2 * d;
  const int img = blockIdx.z;
  const int tyi = blockIdx.x / p.tiles_x;
  const int ty0 = tyi * DW_TH, tx0 = (blockIdx.x - tyi * p.tiles_x) * DW_TW;
  float* tS = dw_smem + ((128u - (dw_smem_u32(dw_smem) & 127u)) & 127u) / 4;
  float* wS = tS + p.tile_floats + (p.res_tma ? DW_TH * DW_TW * DW_CH : 0);
  __shared__ uint64_t bar;
  const int H = p.in.h, W = p.in.w;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dw_smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int res_bytes = p.res_tma ? DW_TH * DW_TW * DW_CH * 4 : 0;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dw_smem_u32(&bar)), "r"(SH * SW * DW_CH * 4 + res_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dw_smem_u32(tS)), "l"(&p.tm[0]), "r"(dw_smem_u32(&bar)), "r"(cin0), "r"(tx0 - d), "r"(ty0 - d), "r"(img)
                 : "memory");
    // the residual tile rides on the same barrier: no residual load (and no DRAM round trip) is left in the row loop
    if (p.res_tma)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(dw_smem_u32(tS + p.tile_floats)), "l"(&p.tm[1]), "r"(dw_smem_u32(&bar)), "r"(cout0), "r"(tx0), "r"(ty0), "r"(img)
                   : "memory");
  }
  // thread = one channel PAIR of one pixel column, every second row of the tile: the nine tap pairs then fit in registers
  // (18) at two 512-thread CTAs per SM, so the only shared-memory reads of the row loop are the nine tile taps. (With
  // channel quads the 36 tap registers did not fit next to three CTAs per SM, and re-reading them from shared memory per
  // row doubled the load wavefronts: the kernel sat at 87 % of the L1/shared data pipe.)
  const int q = tid & 7, lx = (tid >> 3) & 31, ly0 = tid >> 8;
  const int ox = tx0 + lx;
  const bool live = ox < W && wofs + q * 2 < B.c;
  const int c2 = cout0 + q * 2;
  // per-image bases (64-bit once), then 32-bit element offsets inside the image
  const float* resI = p.res.p ? p.res.p + (size_t)img * p.res.h * p.res.w * p.res.ld : nullptr;
  float* outI = p.out.p + (size_t)img * p.out.h * p.out.w * p.out.ld;
  const float* amI = p.amod.p + (size_t)img * p.amod.h * p.amod.w * p.amod.ld;
  const int rows_here = min(DW_TH, H - ty0);
  if (live && resI && !p.res_tma && !(q & 3)) {        // one request per 32-byte sector
#pragma unroll
    for (int r = ly0; r < DW_TH; r += 2)
      if (r < rows_here) asm volatile("prefetch.global.L2 [%0];" ::"l"(resI + ((ty0 + r) * p.res.w + ox) * p.res.ld + c2));
  }
  // taps with the BatchNorm scale folded in: BN(dw(x)) = sum (scale * w) x + shift
  for (int i = tid; i < 9 * DW_CH; i += SA_THREADS) {
    const int c = wofs + (i & 15);
    wS[i] = c < B.c ? __ldg(B.w + (i >> 4) * B.c + c) * __ldg(B.scale + c) : 0.f;
  }
  __syncthreads();
  dw_mbar_wait(&bar, 0);
  if (!live) return;
  f32x2 w[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) w[t] = *reinterpret_cast<const f32x2*>(wS + t * DW_CH + q * 2);
  const f32x2 sh = *reinterpret_cast<const f32x2*>(B.shift + wofs + q * 2);
  const int row_f = SW * DW_CH;
  const int dyf = d * row_f, dxf = d * DW_CH;
  const int vh = H / p.amod.h;
  const int am_x = (ox / (W / p.amod.w)) * p.amod.ld + c2;
  const int am_row = p.amod.w * p.amod.ld;
  const float w0 = p.sa_w0, w1 = p.sa_w1;
  const int res_step = 2 * p.res.w * p.res.ld, out_step = 2 * p.out.w * p.out.ld;
  int res_o = ((ty0 + ly0) * p.res.w + ox) * p.res.ld + c2;
  int out_o = ((ty0 + ly0) * p.out.w + ox) * p.out.ld + c2;
  const bool do16 = p.o16 && c2 < p.o16_c && !(c2 >= p.o16_skip_lo && c2 < p.o16_skip_hi);
  __half* o16I = p.o16 + (size_t)img * p.out.h * p.out.w * p.o16_ld;
  int o16_o = ((ty0 + ly0) * p.out.w + ox) * p.o16_ld + c2;
  const int o16_step = 2 * p.out.w * p.o16_ld;
  const float* base = tS + ly0 * row_f + lx * DW_CH + q * 2;
  // view row of the current output row, stepped without a division per row
  int ay = (ty0 + ly0) / vh, arem = (ty0 + ly0) - ay * vh;
  const float* resS = tS + p.tile_floats + (ly0 * DW_TW + lx) * DW_CH + q * 2;      // staged residual tile [16][32][16]
  const bool res_g = resI && !p.res_tma;
  float2 rn = make_float2(0.f, 0.f), an = make_float2(0.f, 0.f);
  if (ly0 < rows_here) {
    if (res_g) rn = *reinterpret_cast<const float2*>(resI + res_o);
    an = __ldg(reinterpret_cast<const float2*>(amI + ay * am_row + am_x));
  }
#pragma unroll 1
  for (int r = ly0; r < rows_here; r += 2) {
    float2 rc = rn;
    const float2 ac = an;
    if (p.res_tma) rc = *reinterpret_cast<const float2*>(resS + (r - ly0) * DW_TW * DW_CH);
    if (r + 2 < rows_here) {                 // next row's residual and modulation: in flight across this row's arithmetic
      if (res_g) rn = *reinterpret_cast<const float2*>(resI + res_o + res_step);
      arem += 2;
      while (arem >= vh) { arem -= vh; ++ay; }
      an = __ldg(reinterpret_cast<const float2*>(amI + ay * am_row + am_x));
    }
    f32x2 acc = sh;
    float2 xc;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const f32x2 v = *reinterpret_cast<const f32x2*>(base + ky * dyf + kx * dxf);
        if (ky == 1 && kx == 1) unpack2(v, xc.x, xc.y);
        acc = fma2(v, w[ky * 3 + kx], acc);
      }
    float2 a;
    unpack2(acc, a.x, a.y);
    a.x = fmaf(xc.x, fmaf(w0, sa_sigmoid(a.x), w1 * ac.x), rc.x);
    a.y = fmaf(xc.y, fmaf(w0, sa_sigmoid(a.y), w1 * ac.y), rc.y);
    *reinterpret_cast<float2*>(outI + out_o) = a;
    if (do16) *reinterpret_cast<__half2*>(o16I + o16_o) = __floats2half2_rn(a.x, a.y);
    base += 2 * row_f; res_o += res_step; out_o += out_step; o16_o += o16_step;
  }
}

// The same tail as a PERSISTENT kernel: one 512-thread CTA per SM walks the (image, channel chunk, tile) list with two tile
// buffers, so the TMA loads of tile i+1 (input tile with halo + residual tile) are in flight while tile i is computed. In
// sa_tile_kernel 30 % of the warp stall samples sit on the barrier of the tile load: two single-buffered CTAs per SM cannot
// cover a ~2 us load in front of ~2 us of arithmetic.
constexpr int SA_PTHREADS = 512;        // (1024 threads = four row phases per column: 0.346 vs 0.294 ms - the per-tile setup is paid per thread)
__global__ void __launch_bounds__(SA_PTHREADS, 1)
sa_tile_persist_kernel(const __grid_constant__ DwParams p, int tiles_xy, int items, int total) {
  extern __shared__ __align__(16) float dw_smem[];
  const int tid = threadIdx.x;
  const DwBranch& B = p.br[0];
  const int d = B.dh;
  const int SH = DW_TH + 2 * d, SW = DW_TW + 2 * d;
  constexpr int RES_F = DW_TH * DW_TW * DW_CH;
  const int buf_f = p.tile_floats + (p.res_tma ? RES_F : 0);
  float* tS0 = dw_smem + ((128u - (dw_smem_u32(dw_smem) & 127u)) & 127u) / 4;
  float* wS = tS0 + 2 * buf_f;                         // [items][9][16] taps with the BatchNorm scale folded in
  __shared__ uint64_t bars[2];
  const int H = p.in.h, W = p.in.w;
  const uint32_t tx_bytes = (uint32_t)(SH * SW * DW_CH * 4 + (p.res_tma ? RES_F * 4 : 0));
  auto issue = [&](int t, int b) {                     // one thread: both tensor loads of tile t into buffer b
    const int xy = t % tiles_xy; const int r_ = t / tiles_xy;
    const int chunk = r_ % items, img = r_ / items;
    const int tyi = xy / p.tiles_x;
    const int ty0 = tyi * DW_TH, tx0 = (xy - tyi * p.tiles_x) * DW_TW;
    float* dst = tS0 + b * buf_f;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dw_smem_u32(&bars[b])), "r"(tx_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dw_smem_u32(dst)), "l"(&p.tm[0]), "r"(dw_smem_u32(&bars[b])), "r"(B.in_c0 + chunk * DW_CH), "r"(tx0 - d),
                   "r"(ty0 - d), "r"(img) : "memory");
    if (p.res_tma)
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(dw_smem_u32(dst + p.tile_floats)), "l"(&p.tm[1]), "r"(dw_smem_u32(&bars[b])), "r"(B.out_c0 + chunk * DW_CH),
                     "r"(tx0), "r"(ty0), "r"(img) : "memory");
  };
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dw_smem_u32(&bars[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dw_smem_u32(&bars[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if ((int)blockIdx.x < total) issue((int)blockIdx.x, 0);
  }
  for (int i = tid; i < items * 9 * DW_CH; i += SA_PTHREADS) {
    const int c = (i / (9 * DW_CH)) * DW_CH + (i & 15), tap = (i >> 4) % 9;
    wS[i] = c < B.c ? __ldg(B.w + tap * B.c + c) * __ldg(B.scale + c) : 0.f;
  }
  __syncthreads();
  const int q = tid & 7, lx = (tid >> 3) & 31, ly0 = tid >> 8;
  const int row_f = SW * DW_CH;
  const int dyf = d * row_f, dxf = d * DW_CH;
  const int vh = H / p.amod.h, vw = W / p.amod.w;
  const int am_row = p.amod.w * p.amod.ld;
  const float w0 = p.sa_w0, w1 = p.sa_w1;
  constexpr int RS = SA_PTHREADS / 256;            // row phases
  const int out_step = RS * p.out.w * p.out.ld, o16_step = RS * p.out.w * p.o16_ld;
  int it = 0;
  for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
    const int b = it & 1;
    if (tid == 0 && t + (int)gridDim.x < total) issue(t + (int)gridDim.x, b ^ 1);
    const int xy = t % tiles_xy; const int r_ = t / tiles_xy;
    const int chunk = r_ % items, img = r_ / items;
    const int tyi = xy / p.tiles_x;
    const int ty0 = tyi * DW_TH, tx0 = (xy - tyi * p.tiles_x) * DW_TW;
    const int wofs = chunk * DW_CH;
    const int ox = tx0 + lx;
    const bool live = ox < W && wofs + q * 2 < B.c;
    const int c2 = B.out_c0 + wofs + q * 2;
    const float* tS = tS0 + b * buf_f;
    dw_mbar_wait(&bars[b], (uint32_t)(it >> 1) & 1u);
    if (live) {
      f32x2 w[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) w[k] = *reinterpret_cast<const f32x2*>(wS + (chunk * 9 + k) * DW_CH + q * 2);
      const f32x2 sh = *reinterpret_cast<const f32x2*>(B.shift + wofs + q * 2);
      float* outI = p.out.p + (size_t)img * p.out.h * p.out.w * p.out.ld;
      const float* amI = p.amod.p + (size_t)img * p.amod.h * p.amod.w * p.amod.ld + (ox / vw) * p.amod.ld + c2;
      const int rows_here = min(DW_TH, H - ty0);
      int out_o = ((ty0 + ly0) * p.out.w + ox) * p.out.ld + c2;
      const bool do16 = p.o16 && c2 < p.o16_c && !(c2 >= p.o16_skip_lo && c2 < p.o16_skip_hi);
      __half* o16I = p.o16 + (size_t)img * p.out.h * p.out.w * p.o16_ld;
      int o16_o = ((ty0 + ly0) * p.out.w + ox) * p.o16_ld + c2;
      const float* base = tS + ly0 * row_f + lx * DW_CH + q * 2;
      const float* resS = tS + p.tile_floats + (ly0 * DW_TW + lx) * DW_CH + q * 2;
      int ay = (ty0 + ly0) / vh, arem = (ty0 + ly0) - ay * vh;
#pragma unroll 2
      for (int r = ly0; r < rows_here; r += RS) {
        const float2 ac = __ldg(reinterpret_cast<const float2*>(amI + ay * am_row));
        float2 rc = make_float2(0.f, 0.f);
        if (p.res_tma) rc = *reinterpret_cast<const float2*>(resS);
        f32x2 acc = sh;
        float2 xc;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const f32x2 v = *reinterpret_cast<const f32x2*>(base + ky * dyf + kx * dxf);
            if (ky == 1 && kx == 1) unpack2(v, xc.x, xc.y);
            acc = fma2(v, w[ky * 3 + kx], acc);
          }
        float2 a;
        unpack2(acc, a.x, a.y);
        a.x = fmaf(xc.x, fmaf(w0, sa_sigmoid(a.x), w1 * ac.x), rc.x);
        a.y = fmaf(xc.y, fmaf(w0, sa_sigmoid(a.y), w1 * ac.y), rc.y);
        *reinterpret_cast<float2*>(outI + out_o) = a;
        if (do16) *reinterpret_cast<__half2*>(o16I + o16_o) = __floats2half2_rn(a.x, a.y);
        base += RS * row_f; resS += RS * DW_TW * DW_CH; out_o += out_step; o16_o += o16_step;
        arem += RS;
        while (arem >= vh) { arem -= vh; ++ay; }
      }
    }
    __syncthreads();          // buffer b is free for the load issued at the top of the next iteration
  }
}

typedef CUresult (*DwEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DwEncodeFn dw_get_encode() {
  static DwEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<DwEncodeFn>(p);
  });
  return fn;
}

static bool dw_tile_ok(const lfsr_tensor* in, const lfsr_tensor* out, const lfsr_dw_branch* br, int nbr) {
  if (nbr > DW_MAXB) return false;
  if (((uintptr_t)in->ptr & 15) || ((uintptr_t)out->ptr & 15) || (in->ld & 3) || (out->ld & 3)) return false;
  for (int i = 0; i < nbr; ++i) {
    const lfsr_dw_branch& b = br[i];
    if ((b.c & 3) || (b.in_c0 & 3) || (b.out_c0 & 3)) return false;
    if (b.scale && (((uintptr_t)b.scale & 15) || ((uintptr_t)b.shift & 15))) return false;
    const int hy = (b.kh / 2) * b.dil_h, hx = (b.kw / 2) * b.dil_w;
    const size_t bytes = ((size_t)(DW_TH + 2 * hy) * (DW_TW + 2 * hx) * DW_CH + (size_t)b.kh * b.kw * DW_CH) * 4 + 128;
    if (bytes > 200 * 1024) return false;
  }
  return true;
}

}  // namespace lfsr

using namespace lfsr;

// shared by lfsr_dwconv_multi and lfsr_sa_modulate (sa != null: the SA-modulator tail on a single branch)
struct SaTail { float w0, w1; const lfsr_tensor* amod; const lfsr_tensor* res; const lfsr_tensor* out16; int skip_lo, skip_hi; };
static int dw_launch(const lfsr_tensor* in, const lfsr_tensor* out, const lfsr_dw_branch* br, int nbr, const SaTail* sa,
                     bool* used_tile, void* stream) {
  LFSR_REQUIRE(tensor_ok(in) && tensor_ok(out) && br && nbr > 0, "lfsr_dwconv_multi: null/invalid argument");
  LFSR_REQUIRE(in->n == out->n && in->h == out->h && in->w == out->w, "lfsr_dwconv_multi: in/out geometry mismatch");
  LFSR_REQUIRE(in->n <= 65535 && (long long)in->h * in->w * in->ld < 0x7fffffffLL &&
                   (long long)out->h * out->w * out->ld < 0x7fffffffLL, "lfsr_dwconv_multi: tensor too large");
  for (int i = 0; i < nbr; ++i) {
    const lfsr_dw_branch& b = br[i];
    LFSR_REQUIRE(b.w && b.kh > 0 && b.kw > 0 && (b.kh & 1) && (b.kw & 1) && b.dil_h > 0 && b.dil_w > 0,
                 "lfsr_dwconv_multi: branch %d: odd kernels only", i);
    LFSR_REQUIRE(b.c > 0 && b.in_c0 >= 0 && b.out_c0 >= 0 && b.in_c0 + b.c <= in->c && b.out_c0 + b.c <= out->c,
                 "lfsr_dwconv_multi: branch %d: channel window [%d,+%d) -> [%d,+%d) outside %d / %d channels", i, b.in_c0,
                 b.c, b.out_c0, b.c, in->c, out->c);
    LFSR_REQUIRE((b.scale == nullptr) == (b.shift == nullptr), "lfsr_dwconv_multi: scale/shift must come together");
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (used_tile) *used_tile = dw_tile_ok(in, out, br, nbr);
  if (!dw_tile_ok(in, out, br, nbr)) {
    if (sa) return LFSR_OK;       // the caller falls back to its own kernel
    for (int i = 0; i < nbr; ++i) {
      const lfsr_dw_branch& b = br[i];
      TView vi = view_of(in), vo = view_of(out);
      vi.p += b.in_c0; vi.c = b.c;
      vo.p += b.out_c0; vo.c = b.c;
      dim3 blocks(ceil_div(in->h * in->w * b.c, 256), in->n);
      dwconv_kernel<<<blocks, 256, 0, st>>>(vi, vo, b.w, b.scale, b.shift, b.kh, b.kw, b.dil_h, b.dil_w, b.act, b.act_slope);
      int rc = check_launch("dwconv_kernel");
      if (rc != LFSR_OK) return rc;
    }
    return LFSR_OK;
  }
  DwParams p;
  p.in = view_of(in); p.out = view_of(out);
  p.nbr = nbr;
  p.sa = sa ? 1 : 0;
  p.sa_w0 = sa ? sa->w0 : 0.f; p.sa_w1 = sa ? sa->w1 : 0.f;
  p.amod = sa ? view_of(sa->amod) : null_view();
  p.res = sa && sa->res && sa->res->ptr ? view_of(sa->res) : null_view();
  p.o16 = nullptr; p.o16_ld = 0; p.o16_c = 0; p.o16_skip_lo = p.o16_skip_hi = 0;
  if (sa && sa->out16 && sa->out16->ptr) { p.o16 = (__half*)sa->out16->ptr; p.o16_ld = (int)sa->out16->ld; p.o16_c = sa->out16->c; p.o16_skip_lo = sa->skip_lo; p.o16_skip_hi = sa->skip_hi; }
  int items = 0, w_floats = 0;
  size_t tile_floats = 0;
  for (int i = 0; i < nbr; ++i) {
    const lfsr_dw_branch& b = br[i];
    DwBranch& d = p.br[i];
    d.w = b.w; d.scale = b.scale; d.shift = b.shift;
    d.kh = b.kh; d.kw = b.kw; d.dh = b.dil_h; d.dw = b.dil_w;
    d.in_c0 = b.in_c0; d.out_c0 = b.out_c0; d.c = b.c; d.act = b.act; d.slope = b.act_slope;
    d.item0 = items;
    items += ceil_div(b.c, DW_CH);
    const int hy = (b.kh / 2) * b.dil_h, hx = (b.kw / 2) * b.dil_w;
    const size_t tf = (size_t)(DW_TH + 2 * hy) * (DW_TW + 2 * hx) * DW_CH;
    if (tf > tile_floats) tile_floats = tf;
    if (b.kh * b.kw * DW_CH > w_floats) w_floats = b.kh * b.kw * DW_CH;
  }
  p.share = 0; p.share_b = 0; p.res_tma = 0;
  if (!sa && nbr > 1 && !dbg_env("LFSR_DW_NO_SHARE")) {
    bool same = true;
    size_t best = 0;
    int wsum = 0;
    for (int i = 0; i < nbr; ++i) {
      same = same && br[i].in_c0 == br[0].in_c0 && br[i].c == br[0].c;
      const size_t tf = (size_t)(DW_TH + 2 * (br[i].kh / 2) * br[i].dil_h) * (DW_TW + 2 * (br[i].kw / 2) * br[i].dil_w);
      if (tf > best) { best = tf; p.share_b = i; }
      wsum += br[i].kh * br[i].kw * DW_CH;
    }
    // the widest branch must cover the others in both directions
    for (int i = 0; i < nbr && same; ++i)
      same = (br[i].kh / 2) * br[i].dil_h <= (br[p.share_b].kh / 2) * br[p.share_b].dil_h &&
             (br[i].kw / 2) * br[i].dil_w <= (br[p.share_b].kw / 2) * br[p.share_b].dil_w;
    if (same && (tile_floats + (size_t)wsum) * sizeof(float) + 128 <= 200 * 1024) {
      p.share = 1;
      w_floats = wsum;
      items = ceil_div(br[0].c, DW_CH);
    }
  }
  LFSR_REQUIRE(items <= 65535, "lfsr_dwconv_multi: too many channel chunks");
  p.tiles_x = ceil_div(in->w, DW_TW);
  p.w_floats = w_floats;
  p.tile_floats = (int)tile_floats;
  static const bool no_tma = dbg_env("LFSR_DW_NO_TMA") != nullptr;
  DwEncodeFn encode = no_tma ? nullptr : dw_get_encode();
  p.use_tma = encode != nullptr;
  for (int i = 0; i < nbr && p.use_tma; ++i) {
    const lfsr_dw_branch& b = br[i];
    const int hy = (b.kh / 2) * b.dil_h, hx = (b.kw / 2) * b.dil_w;
    const cuuint64_t ld_b = (cuuint64_t)in->ld * 4;
    cuuint64_t dims[4] = {(cuuint64_t)in->c, (cuuint64_t)in->w, (cuuint64_t)in->h, (cuuint64_t)in->n};
    cuuint64_t strides[3] = {ld_b, ld_b * in->w, ld_b * in->w * in->h};
    cuuint32_t box[4] = {(cuuint32_t)DW_CH, (cuuint32_t)(DW_TW + 2 * hx), (cuuint32_t)(DW_TH + 2 * hy), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (box[1] > 256 || box[2] > 256 ||
        encode(&p.tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, in->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      p.use_tma = 0;
  }
  const size_t smem = (tile_floats + (size_t)w_floats) * sizeof(float) + 128;
  static DevOnce once;
  if (smem > 48 * 1024 && once.need()) {
    if (opt_in_smem(dw_tile_kernel<false>, 200 * 1024 + 4096, "lfsr_dwconv_multi") ||
        opt_in_smem(dw_tile_kernel<true>, 200 * 1024 + 4096, "lfsr_dwconv_multi")) return LFSR_ERR_CUDA;
    once.done();
  }
  dim3 grid(p.tiles_x * ceil_div(in->h, DW_TH), items, in->n);
  if (sa && p.use_tma && nbr == 1 && br[0].kh == 3 && br[0].kw == 3 && br[0].dil_h == br[0].dil_w && br[0].scale &&
      br[0].act == LFSR_ACT_SIGMOID && in->h % p.amod.h == 0 && in->w % p.amod.w == 0 &&
      (!p.res.p || (p.res.h == in->h && p.res.w == in->w))) {
    size_t smem_sa = smem;
    if (p.res.p && smem + DW_TH * DW_TW * DW_CH * 4 <= 110 * 1024) {        // residual tile by TMA (still two CTAs per SM)
      const lfsr_tensor* rt = sa->res;
      const cuuint64_t ld_b = (cuuint64_t)rt->ld * 4;
      cuuint64_t dims[4] = {(cuuint64_t)rt->c, (cuuint64_t)rt->w, (cuuint64_t)rt->h, (cuuint64_t)rt->n};
      cuuint64_t strides[3] = {ld_b, ld_b * rt->w, ld_b * rt->w * rt->h};
      cuuint32_t box[4] = {(cuuint32_t)DW_CH, (cuuint32_t)DW_TW, (cuuint32_t)DW_TH, 1};
      cuuint32_t estr[4] = {1, 1, 1, 1};
      if (encode(&p.tm[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rt->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
        p.res_tma = 1;
        smem_sa += DW_TH * DW_TW * DW_CH * 4;
      }
    }
    const int items_sa = ceil_div(br[0].c, DW_CH);
    const size_t smem_p = 2 * ((size_t)tile_floats + (p.res_tma ? DW_TH * DW_TW * DW_CH : 0)) * 4 + (size_t)items_sa * 9 * DW_CH * 4 + 128;
    if ((p.res_tma || !p.res.p) && smem_p <= 225 * 1024 && !dbg_env("LFSR_SA_NO_PERSIST")) {
      static DevOnce once_sp;
      if (once_sp.need()) {
        if (opt_in_smem(sa_tile_persist_kernel, 225 * 1024, "lfsr_sa_modulate")) return LFSR_ERR_CUDA;
        once_sp.done();
      }
      const int tiles_xy = p.tiles_x * ceil_div(in->h, DW_TH);
      const long long total = (long long)tiles_xy * items_sa * in->n;
      if (total <= 0x7fffffffLL) {
        const int nsm = sm_count_current();
        const int g = total < nsm ? (int)total : nsm;
        sa_tile_persist_kernel<<<g, SA_PTHREADS, smem_p, st>>>(p, tiles_xy, items_sa, (int)total);
        return check_launch("sa_tile_persist_kernel");
      }
    }
    static DevOnce once_sa;
    if (once_sa.need()) {
      if (opt_in_smem(sa_tile_kernel, 200 * 1024 + 4096, "lfsr_sa_modulate")) return LFSR_ERR_CUDA;
      once_sa.done();
    }
    sa_tile_kernel<<<grid, SA_THREADS, smem_sa, st>>>(p);
    return check_launch("sa_tile_kernel");
  }
  if (p.share) dw_tile_kernel<true><<<grid, 256, smem, st>>>(p);
  else dw_tile_kernel<false><<<grid, 256, smem, st>>>(p);
  return check_launch("dw_tile_kernel");
}

extern "C" int lfsr_dwconv_multi(const lfsr_tensor* in, const lfsr_tensor* out, const lfsr_dw_branch* br, int nbr,
                                 void* stream) {
  return dw_launch(in, out, br, nbr, nullptr, nullptr, stream);
}

// SA-modulator tail on the tiled depthwise kernel; *handled = 0 when the tensors do not qualify (caller falls back)
int lfsr_sa_modulate_tiled(const lfsr_tensor* x, const float* dw_w, const float* bn_scale, const float* bn_shift,
                           const lfsr_tensor* amod, float w0, float w1, const lfsr_tensor* res, const lfsr_tensor* out, int dil,
                           int* handled, void* stream, const lfsr_tensor* out16, int skip_lo, int skip_hi) {
  lfsr_dw_branch b;
  b.w = dw_w; b.scale = bn_scale; b.shift = bn_shift;
  b.kh = 3; b.kw = 3; b.dil_h = dil; b.dil_w = dil;
  b.in_c0 = 0; b.out_c0 = 0; b.c = x->c;
  b.act = LFSR_ACT_SIGMOID; b.act_slope = 0.f;
  auto al = [](const lfsr_tensor* t) { return t && t->ptr && t->ld % 4 == 0 && (((uintptr_t)t->ptr) & 15) == 0; };
  *handled = 0;
  if (!al(amod) || (res && res->ptr && !al(res))) return LFSR_OK;
  SaTail sa{w0, w1, amod, res, out16, skip_lo, skip_hi};
  bool used = false;
  int rc = dw_launch(x, out, &b, 1, &sa, &used, stream);
  *handled = used ? 1 : 0;
  return rc;
}

extern "C" int lfsr_dwconv_f32(const lfsr_tensor* in, const float* w_packed, const float* scale, const float* shift,
                               const lfsr_tensor* out, int kh, int kw, int dil_h, int dil_w, int act, float act_slope,
                               void* stream) {
  LFSR_REQUIRE(tensor_ok(in) && tensor_ok(out) && w_packed, "lfsr_dwconv_f32: null/invalid tensor");
  LFSR_REQUIRE(in->c == out->c, "lfsr_dwconv_f32: in/out shape mismatch");
  lfsr_dw_branch b;
  b.w = w_packed; b.scale = scale; b.shift = shift;
  b.kh = kh; b.kw = kw; b.dil_h = dil_h; b.dil_w = dil_w;
  b.in_c0 = 0; b.out_c0 = 0; b.c = in->c;
  b.act = act; b.act_slope = act_slope;
  return lfsr_dwconv_multi(in, out, &b, 1, stream);
}
