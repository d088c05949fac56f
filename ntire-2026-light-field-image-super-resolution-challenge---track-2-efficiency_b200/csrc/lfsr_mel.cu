// HBM/L1-bound kernels for the thin layers of the path:
//   * conv_small_cout: KxK conv with 1..4 output channels (the 54->1 / 64->1 reconstruction heads,
//     MyEfficientLFNet.py:70-73, EPIT.py:48): one thread per output pixel, channel-vectorised loads,
//     weights broadcast from shared memory, residual add fused (in place on the interpolation skip).
//   * mel_epi_branch: the whole MultiScaleEPIBlock (MyEfficientLFNet.py:278-327) in one pass:
//     dw 1xK / Kx1 / 3x3-dilated -> 1x1 + LReLU each -> concat -> 1x1 + LReLU. One thread per pixel.
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>
#include <mutex>
#include "lfsr_common.cuh"
#include "lfsr_ptx.cuh"

namespace lfsr {

// ---- small-cout conv -------------------------------------------------------------------------
struct SmallArgs {
  TView in, out, res;
  const float* w;      // [tap][cin][cout]
  const float* bias;
  int kh, kw, dh, dw, ph, pw, cout, act;
  float slope, alpha;
  int tiles_x, tiles_y;
  int vec16;           // channel vectors are whole 16-byte chunks at 16-byte aligned addresses
};

// Block = 32x8 output pixels. The input tile with halo is staged in shared memory with fully coalesced
// loads (consecutive threads read consecutive 8 bytes of a pixel's channel vector); the per-pixel channel
// stride CS satisfies CS % 8 == 4, which makes the 128-bit reads of the compute phase (lane = pixel,
// stride CS floats) bank-conflict free. Weights are broadcast from shared memory.
constexpr int ST_W = 32, ST_H = 8;
template <int COUT>
__global__ void __launch_bounds__(256)
conv_small_cout_kernel(const SmallArgs a, int CS) {
  extern __shared__ __align__(16) float smem[];
  const int C = a.in.c, taps = a.kh * a.kw;
  const int hw = ST_W + a.dw * (a.kw - 1), hh = ST_H + a.dh * (a.kh - 1);
  float* tile = smem;                          // [hh*hw][CS]
  float* wsm = smem + hh * hw * CS;            // [taps][COUT][CS]
  int t_ = blockIdx.x;
  const int tx0 = (t_ % a.tiles_x) * ST_W; t_ /= a.tiles_x;
  const int ty0 = (t_ % a.tiles_y) * ST_H;
  const int img = t_ / a.tiles_y;
  const int C2 = C >> 1, CS2 = CS >> 1;
  // weights: [tap][cin][cout] -> [tap][cout][CS] (zero padded)
  for (int i = threadIdx.x; i < taps * COUT * CS; i += 256) {
    const int c = i % CS, r = i / CS;
    const int o = r % COUT, t = r / COUT;
    wsm[i] = c < C ? __ldg(a.w + ((size_t)t * C + c) * COUT + o) : 0.f;
  }
  // input tile (zero outside the image = conv zero padding). When the channel vector is a whole number of 16-byte
  // chunks it is staged with cp.async (LDGSTS, zero-fill for out-of-image pixels): ~19 independent async copies per
  // thread and no register round trip; the CS padding lanes are cleared separately.
  if (a.vec16) {
    const int chunks = C >> 2;                       // 16-byte chunks per pixel
    const int pad4 = (CS - C) >> 2;
    for (int i = threadIdx.x; i < hh * hw * pad4; i += 256) {
      const int pix = i / pad4, q = i - pix * pad4;
      reinterpret_cast<float4*>(tile + pix * CS + C)[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = threadIdx.x; i < hh * hw * chunks; i += 256) {
      const int ch = i % chunks, pix = i / chunks;
      const int ly = pix / hw, lx = pix - ly * hw;
      const int iy = ty0 - a.ph + ly, ix = tx0 - a.pw + lx;
      const bool inside = iy >= 0 && iy < a.in.h && ix >= 0 && ix < a.in.w;
      const float* src = a.in.p + (inside ? a.in.pix(img, iy, ix) : 0) + ch * 4;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile + pix * CS + ch * 4);
      const int nbytes = inside ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  } else {
    for (int i = threadIdx.x; i < hh * hw * CS2; i += 256) {
      const int c2 = i % CS2, pix = i / CS2;
      const int ly = pix / hw, lx = pix - ly * hw;
      const int iy = ty0 - a.ph + ly, ix = tx0 - a.pw + lx;
      float2 v = make_float2(0.f, 0.f);
      if (c2 < C2 && iy >= 0 && iy < a.in.h && ix >= 0 && ix < a.in.w)
        v = __ldg(reinterpret_cast<const float2*>(a.in.p + a.in.pix(img, iy, ix)) + c2);
      reinterpret_cast<float2*>(tile)[i] = v;
    }
  }
  __syncthreads();
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int ox = tx0 + lx, oy = ty0 + ly;
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
  for (int ky = 0; ky < a.kh; ++ky)
    for (int kx = 0; kx < a.kw; ++kx) {
      const float4* src = reinterpret_cast<const float4*>(tile + ((ly + ky * a.dh) * hw + lx + kx * a.dw) * CS);
      const float4* wt = reinterpret_cast<const float4*>(wsm + (ky * a.kw + kx) * COUT * CS);
#pragma unroll 4
      for (int c4 = 0; c4 < (CS >> 2); ++c4) {
        const float4 v = src[c4];
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
          const float4 w = wt[o * (CS >> 2) + c4];
          acc[o] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[o]))));
        }
      }
    }
  if (ox >= a.out.w || oy >= a.out.h) return;
  const size_t ob = a.out.pix(img, oy, ox);
  const size_t rb = a.res.p ? a.res.pix(img, oy, ox) : 0;
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    float v = acc[o] + (a.bias ? __ldg(a.bias + o) : 0.f);
    v = apply_act(v, a.act, a.slope) * a.alpha;
    if (a.res.p) v += a.res.p[rb + o];
    a.out.p[ob + o] = v;
  }
}

// ---- fused MultiScaleEPIBlock ---------------------------------------------------------------------
constexpr int EC = 18;   // channels of the EPI split (54 - 2*18, MyEfficientLFNet.py:129)
struct EpiArgs {
  TView in, out;
  const float* w;      // packed: dw_h[KL][EC] | dw_v[KL][EC] | dw_d[9][EC] | pw_h[EC][EC] | pw_v | pw_d | fuse[3*EC][EC]
  int KL, dil;
  float slope;
  int tiles_x, tiles_y;
};

__device__ __forceinline__ void load18(const float* src, float* v) {
#pragma unroll
  for (int i = 0; i < EC / 2; ++i) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(src) + i);
    v[2 * i] = t.x; v[2 * i + 1] = t.y;
  }
}

// Block = 32x8 output pixels. The 18-channel input tile with a halo of max(klen/2, dil) pixels is staged with
// cp.async as 20-float pixels (18 + the 2 zero pad lanes of the grouped trunk layout) - 80 B = five 16-byte chunks -
// and every tap is then five conflict-free LDS.128 (stride 20 floats) instead of nine strided 8-byte global loads.
constexpr int EP_W = 32, EP_H = 8, EP_CS = 20;
__device__ __forceinline__ void lds20(const float* src, float* v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = reinterpret_cast<const float4*>(src)[i];
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
  const float2 t = *reinterpret_cast<const float2*>(src + 16);
  v[16] = t.x; v[17] = t.y;
}

// 18 weights of one row (padded to EP_CS = 20 floats, 16-byte aligned) as 9 packed pairs; every lane reads the same
// address, so these are broadcast LDS.128
__device__ __forceinline__ void ldrow9(const float* row, f32x2* w) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = reinterpret_cast<const float4*>(row)[i];
    w[2 * i] = pack2(t.x, t.y); w[2 * i + 1] = pack2(t.z, t.w);
  }
  const float2 t = *reinterpret_cast<const float2*>(row + 16);
  w[8] = pack2(t.x, t.y);
}

// All arithmetic is packed 2 x fp32 FMA (FFMA2) on channel pairs with vector weight loads: ~2100 issued instructions
// per pixel instead of ~5000 with scalar FMAs and scalar weight reads; each half is a plain fma.rn in the same order
// as the scalar formulation, so results are unchanged.
__global__ void __launch_bounds__(256)
mel_epi_branch_kernel(const EpiArgs a) {
  extern __shared__ __align__(16) float sw[];
  const int KL = a.KL;
  const int half = KL / 2;
  const int halo = half > a.dil ? half : a.dil;
  const int hw = EP_W + 2 * halo, hh = EP_H + 2 * halo;
  const int n_rows = 2 * KL + 9 + 6 * EC;               // dw_h | dw_v | dw_d | pw_h pw_v pw_d | fuse, EC floats per row
  float* tile = sw + n_rows * EP_CS;                    // [hh*hw][EP_CS]
  int t_ = blockIdx.x;
  const int tx0 = (t_ % a.tiles_x) * EP_W; t_ /= a.tiles_x;
  const int ty0 = (t_ % a.tiles_y) * EP_H;
  const int img = t_ / a.tiles_y;
  for (int i = threadIdx.x; i < hh * hw * 5; i += 256) {
    const int ch = i % 5, pix = i / 5;
    const int ly = pix / hw, lx = pix - ly * hw;
    const int iy = ty0 - halo + ly, ix = tx0 - halo + lx;
    const bool inside = iy >= 0 && iy < a.in.h && ix >= 0 && ix < a.in.w;
    const float* src = a.in.p + (inside ? a.in.pix(img, iy, ix) : 0) + ch * 4;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tile + pix * EP_CS + ch * 4);
    const int nbytes = inside ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int i = threadIdx.x; i < n_rows * EP_CS; i += 256) {       // rows re-pitched from EC to EP_CS floats
    const int r = i / EP_CS, c = i - r * EP_CS;
    sw[i] = c < EC ? __ldg(a.w + r * EC + c) : 0.f;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const float* dwh = sw;
  const float* dwv = dwh + KL * EP_CS;
  const float* dwd = dwv + KL * EP_CS;
  const float* pw = dwd + 9 * EP_CS;         // 3 x [EC in] rows of EC out
  const float* fu = pw + 3 * EC * EP_CS;     // [3*EC in] rows of EC out
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  const int ox = tx0 + lx, oy = ty0 + ly;
  const float* centre = tile + ((ly + halo) * hw + lx + halo) * EP_CS;
  constexpr int NP = EC / 2;
  f32x2 out2[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) out2[i] = pack2(0.f, 0.f);
#pragma unroll 1
  for (int br = 0; br < 3; ++br) {
    f32x2 t2[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) t2[i] = pack2(0.f, 0.f);
    const int ntap = br == 2 ? 9 : KL;
    const float* dwb = br == 0 ? dwh : (br == 1 ? dwv : dwd);
    for (int k = 0; k < ntap; ++k) {
      int dy = 0, dx = 0;
      if (br == 0) dx = k - half;
      else if (br == 1) dy = k - half;
      else { dy = (k / 3 - 1) * a.dil; dx = (k % 3 - 1) * a.dil; }
      f32x2 v[NP], w[NP];
      ldrow9(centre + (dy * hw + dx) * EP_CS, v);          // out-of-image pixels were zero-filled
      ldrow9(dwb + k * EP_CS, w);
#pragma unroll
      for (int i = 0; i < NP; ++i) t2[i] = fma2(v[i], w[i], t2[i]);
    }
    float t[EC];
#pragma unroll
    for (int i = 0; i < NP; ++i) unpack2(t2[i], t[2 * i], t[2 * i + 1]);
    // pointwise 1x1 (all EC outputs at once, input channel by input channel) + LReLU
    f32x2 s2[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) s2[i] = pack2(0.f, 0.f);
    const float* pwb = pw + br * EC * EP_CS;
#pragma unroll
    for (int c = 0; c < EC; ++c) {
      f32x2 w[NP];
      ldrow9(pwb + c * EP_CS, w);
      const f32x2 tb = pack2(t[c], t[c]);
#pragma unroll
      for (int i = 0; i < NP; ++i) s2[i] = fma2(tb, w[i], s2[i]);
    }
    float sv[EC];
#pragma unroll
    for (int i = 0; i < NP; ++i) unpack2(s2[i], sv[2 * i], sv[2 * i + 1]);
    // ... straight into the fuse 1x1 accumulation
    const float* fub = fu + br * EC * EP_CS;
#pragma unroll
    for (int o = 0; o < EC; ++o) {
      const float so = sv[o] > 0.f ? sv[o] : sv[o] * a.slope;
      f32x2 w[NP];
      ldrow9(fub + o * EP_CS, w);
      const f32x2 sb = pack2(so, so);
#pragma unroll
      for (int i = 0; i < NP; ++i) out2[i] = fma2(sb, w[i], out2[i]);
    }
  }
  if (ox >= a.in.w || oy >= a.in.h) return;
  float* dst = a.out.p + a.out.pix(img, oy, ox);
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float x0, x1;
    unpack2(out2[i], x0, x1);
    x0 = x0 > 0.f ? x0 : x0 * a.slope;
    x1 = x1 > 0.f ? x1 : x1 * a.slope;
    reinterpret_cast<float2*>(dst)[i] = make_float2(x0, x1);
  }
}


// ---- fused MultiScaleEPIBlock, pointwise contractions on the tensor cores -----------------------------------
// Same block as mel_epi_branch_kernel, but only the depthwise taps stay on the CUDA cores. The 1x1 convolutions are two
// tcgen05 GEMMs per 128-pixel row block whose A operands the CTA writes itself:
//   stage 1   T[128 px][54 = 3 x 18 depthwise results, fp16] x blockdiag(pw_h, pw_v, pw_d)  -> TMEM (64 fp32 columns)
//   stage 2   LReLU(stage 1) as fp16, written over T                       x fuse[54 -> 18] -> TMEM (32 columns, reused)
// (K = 54 padded to 64 fp16 = one 128-byte SWIZZLE_128B row per pixel, so each GEMM is four K = 16 MMAs.) The block-
// diagonal zeros triple the MMA work, which is still < 2 % of the tensor pipe; what the kernel saves is the 108 broadcast
// weight rows (540 LDS.128) and 972 FFMA2 per pixel the CUDA-core version spends on the contractions.
// CTA = 256 threads = 32 x 8 pixels = two M = 128 blocks (thread t <-> row t & 127 of block t >> 7 <-> TMEM lane t & 127).
// The A rows alias the staged input tile (dead after the depthwise phase), so 3 CTAs still fit one SM.
namespace et {
using namespace ptx;
constexpr int kA = 16384;            // one M = 128 operand block: 128 rows x 128 B
constexpr int kB1 = 8192, kB2 = 4096;

__device__ __forceinline__ void store_row(uint32_t a_row, uint32_t sw, const uint32_t* pk) {
#pragma unroll
  for (int u = 0; u < 8; ++u)
    st_shared_v4(a_row + ((((uint32_t)u) ^ sw) << 4), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
}
}  // namespace et

__global__ void __launch_bounds__(256, 3)
mel_epi_branch_tc_kernel(const EpiArgs a) {
  using namespace et;
  extern __shared__ uint8_t et_raw[];
  const uint32_t raw = smem_u32(et_raw);
  uint8_t* smem = et_raw + (((raw + 1023u) & ~1023u) - raw);
  const int KL = a.KL, half = KL / 2;
  const int halo = half > a.dil ? half : a.dil;
  const int hw = EP_W + 2 * halo, hh = EP_H + 2 * halo;
  const int n_rows = 2 * KL + 9;                         // depthwise tap rows: dw_h | dw_v | dw_d
  int tile_bytes = hh * hw * EP_CS * 4;
  if (tile_bytes < 2 * kA) tile_bytes = 2 * kA;
  uint8_t* B1 = smem;                                    // [64 out][64 in] fp16, K-major SWIZZLE_128B
  uint8_t* B2 = smem + kB1;                              // [32 out][64 in]
  float* tile = reinterpret_cast<float*>(smem + kB1 + kB2);     // [hh*hw][EP_CS]; later the two A blocks
  float* sw = reinterpret_cast<float*>(smem + kB1 + kB2 + tile_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sw + ((n_rows * EP_CS + 3) & ~3));
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  int t_ = blockIdx.x;
  const int tx0 = (t_ % a.tiles_x) * EP_W; t_ /= a.tiles_x;
  const int ty0 = (t_ % a.tiles_y) * EP_H;
  const int img = t_ / a.tiles_y;
  // input tile with halo (zero filled outside the image)
  for (int i = tid; i < hh * hw * 5; i += 256) {
    const int ch = i % 5, pix = i / 5;
    const int ly = pix / hw, lx = pix - ly * hw;
    const int iy = ty0 - halo + ly, ix = tx0 - halo + lx;
    const bool inside = iy >= 0 && iy < a.in.h && ix >= 0 && ix < a.in.w;
    const float* src = a.in.p + (inside ? a.in.pix(img, iy, ix) : 0) + ch * 4;
    const uint32_t dst = smem_u32(tile + pix * EP_CS + ch * 4);
    const int nbytes = inside ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (tid == 0) {
    mbar_init(bars, 1);
    mbar_init(bars + 1, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(tmem_slot, 128);
  }
  for (int i = tid; i < n_rows * EP_CS; i += 256) {       // tap rows re-pitched from EC to EP_CS floats
    const int r = i / EP_CS, c = i - r * EP_CS;
    sw[i] = c < EC ? __ldg(a.w + r * EC + c) : 0.f;
  }
  // B operands, one 16-byte unit (8 input channels of one output row) at a time
  const float* pwg = a.w + n_rows * EC;                   // 3 x [EC in][EC out]
  const float* fug = pwg + 3 * EC * EC;                   // [3*EC in][EC out]
  for (int i = tid; i < (64 + 32) * 8; i += 256) {
    const int n = i >> 3, u = i & 7;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = 8 * u + j;
      float x = 0.f;
      if (k < 3 * EC) {
        if (n < 64) {
          const int brk = k / EC, c = k - brk * EC;
          if (n < 3 * EC && n / EC == brk) x = __ldg(pwg + brk * EC * EC + c * EC + (n - brk * EC));
        } else if (n - 64 < EC) {
          x = __ldg(fug + k * EC + (n - 64));
        }
      }
      v[j] = x;
    }
    const int row = n < 64 ? n : n - 64;
    const uint32_t base = smem_u32(n < 64 ? B1 : B2) + (uint32_t)row * 128u;
    st_shared_v4(base + ((((uint32_t)u) ^ ((uint32_t)row & 7u)) << 4), pack_f16x2(v[0], v[1]), pack_f16x2(v[2], v[3]),
                 pack_f16x2(v[4], v[5]), pack_f16x2(v[6], v[7]));
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float* dwh = sw;
  const float* dwv = dwh + KL * EP_CS;
  const float* dwd = dwv + KL * EP_CS;
  const int lx = tid & 31, ly = tid >> 5;
  const int ox = tx0 + lx, oy = ty0 + ly;
  const float* centre = tile + ((ly + halo) * hw + lx + halo) * EP_CS;
  constexpr int NP = EC / 2;
  uint32_t pk[32];                                        // this pixel's A row: 54 fp16 + 10 zeros
#pragma unroll
  for (int i = 3 * NP; i < 32; ++i) pk[i] = 0u;
#pragma unroll
  for (int br = 0; br < 3; ++br) {
    f32x2 t2[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) t2[i] = pack2(0.f, 0.f);
    const int ntap = br == 2 ? 9 : KL;
    const float* dwb = br == 0 ? dwh : (br == 1 ? dwv : dwd);
#pragma unroll 1
    for (int k = 0; k < ntap; ++k) {
      int dy = 0, dx = 0;
      if (br == 0) dx = k - half;
      else if (br == 1) dy = k - half;
      else { dy = (k / 3 - 1) * a.dil; dx = (k % 3 - 1) * a.dil; }
      f32x2 v[NP], w[NP];
      ldrow9(centre + (dy * hw + dx) * EP_CS, v);          // out-of-image pixels were zero-filled
      ldrow9(dwb + k * EP_CS, w);
#pragma unroll
      for (int i = 0; i < NP; ++i) t2[i] = fma2(v[i], w[i], t2[i]);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      float x0, x1;
      unpack2(t2[i], x0, x1);
      pk[br * NP + i] = pack_f16x2(x0, x1);
    }
  }
  __syncthreads();                                         // every thread is done with the input tile
  const uint32_t a_blk = smem_u32(tile) + (uint32_t)(tid >> 7) * kA;
  const uint32_t a_row = a_blk + (uint32_t)(tid & 127) * 128u;
  const uint32_t swz = (uint32_t)tid & 7u;
  store_row(a_row, swz, pk);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      tc_fence_after();
      const uint32_t id1 = make_idesc(0, 64);
      const uint64_t db = make_smem_desc(smem_u32(B1));
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const uint64_t da = make_smem_desc(smem_u32(tile) + m * kA);
        umma_f16<0>(tmem + m * 64, da, db, id1);
        umma_f16<1>(tmem + m * 64, da + 2, db + 2, id1);
        umma_f16<1>(tmem + m * 64, da + 4, db + 4, id1);
        umma_f16<1>(tmem + m * 64, da + 6, db + 6, id1);
      }
      umma_commit(bars);
    }
    __syncwarp();
  }
  mbar_wait(bars, 0);
  tc_fence_after();
  const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tid >> 7) * 64u;
  {
    float s[32];
#pragma unroll
    for (int hlf = 0; hlf < 2; ++hlf) {
      tmem_ld32(tlane + hlf * 32, s);
      tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float x0 = s[2 * i], x1 = s[2 * i + 1];
        x0 = x0 > 0.f ? x0 : x0 * a.slope;
        x1 = x1 > 0.f ? x1 : x1 * a.slope;
        pk[hlf * 16 + i] = pack_f16x2(x0, x1);
      }
    }
  }
  store_row(a_row, swz, pk);                               // stage-1 MMAs have retired: the row is rewritten in place
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      tc_fence_after();
      const uint32_t id2 = make_idesc(0, 32);
      const uint64_t db = make_smem_desc(smem_u32(B2));
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        const uint64_t da = make_smem_desc(smem_u32(tile) + m * kA);
        umma_f16<0>(tmem + m * 64, da, db, id2);
        umma_f16<1>(tmem + m * 64, da + 2, db + 2, id2);
        umma_f16<1>(tmem + m * 64, da + 4, db + 4, id2);
        umma_f16<1>(tmem + m * 64, da + 6, db + 6, id2);
      }
      umma_commit(bars + 1);
    }
    __syncwarp();
  }
  mbar_wait(bars + 1, 0);
  tc_fence_after();
  {
    float o[32];
    tmem_ld32(tlane, o);
    tmem_wait_ld();
    if (ox < a.in.w && oy < a.in.h) {
      float* dst = a.out.p + a.out.pix(img, oy, ox);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        float x0 = o[2 * i], x1 = o[2 * i + 1];
        x0 = x0 > 0.f ? x0 : x0 * a.slope;
        x1 = x1 > 0.f ? x1 : x1 * a.slope;
        reinterpret_cast<float2*>(dst)[i] = make_float2(x0, x1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}


// ---- fused MultiScaleEPIBlock, depthwise taps AND pointwise contractions on the tensor cores -------------------------
// dw_b followed by pw_b is linear, so stage 1 of the block is a sum over the 31 taps of
//     X[pixel + tap shift][c] . (dw_b[tap][c] * pw_b[c][o])
// i.e. one tcgen05 MMA (M = 128 pixels, N = 32, K = 16 channels) per tap whose A operand is the SAME shared-memory tile read
// at a shifted row: the input tile with halo is staged once as fp16 pixels of 32 bytes (channels 0..15, K-major
// SWIZZLE_32B) and a tap shift is just a descriptor start address (the swizzle is a function of the absolute address).
// MMA rows are consecutive positions of the PADDED tile (pitch P = 32 + 2 halo), so ~25 % of them are halo columns whose
// results are dropped. Channels 16 and 17 do not fit the K = 16 instruction: their three depthwise sums are computed on
// the CUDA cores (31 FFMA2 per pixel) and enter stage 1 through one extra K = 16 MMA. Stage 2 (LReLU, fuse 1x1, LReLU) is
// the same as in mel_epi_branch_tc_kernel. All B operands come pre-swizzled from lfsr_mel_epi_pack (one bulk copy).
// Per pixel the CUDA cores now do ~90 shared-memory wavefronts instead of ~925; the tensor pipe reads 4 KB per tap and
// 128-row block.
namespace em {
using namespace ptx;
constexpr int TW = 32;               // output tile width
constexpr int kTapB = 1024;          // one tap's B operand: [32 out][16 in] fp16, SWIZZLE_32B
constexpr int kBex = 96 * 32;        // extra-channel B operand: [96 out][16 in]
constexpr int kB2 = 4096;            // fuse B operand: [32 out][64 in], SWIZZLE_128B
constexpr int kA2 = 16384;           // stage-2 A block: [128][64] fp16, SWIZZLE_128B
constexpr int kAex = 4096;           // extra-channel A block: [128][16] fp16, SWIZZLE_32B
constexpr int kMaxTaps = 40;

// K-major SWIZZLE_32B operand: rows of 32 bytes, 8-row groups 256 bytes apart; any 32-byte aligned start
__device__ __forceinline__ uint64_t desc32(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(256 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)6 << 61);
}
// byte address of 16-byte unit u (0/1) of the 32-byte row starting at absolute shared address `row`
__device__ __forceinline__ uint32_t unit32(uint32_t row, uint32_t u) { return row + ((u ^ ((row >> 7) & 1u)) << 4); }

__host__ __device__ inline int b2_offset(int ntap) { const int o = ntap * kTapB + kBex; return o; }
__host__ __device__ inline int image_bytes(int ntap) { return ntap * kTapB + kBex + kB2 + ((ntap * 8 + 255) & ~255); }
}  // namespace em

struct EpiMmaArgs {
  TView in, out;
  const uint8_t* packed;     // lfsr_mel_epi_pack image (device)
  int KL, dil, halo, ntap;
  int P, G0, R, npx;         // padded pitch, first MMA row (flattened padded position), output rows per tile, staged pixels
  float slope;
  int tiles_x, tiles_y, total_tiles;
};

// Persistent, one CTA per SM, 14 warps: 0..7 = epilogue workers (thread t <-> MMA row t & 127 of block t >> 7), 8 = MMA issuer,
// 9 = TMA producer, 10..13 = the CUDA-core side channel for channels 16, 17 (runs a tile ahead of the epilogues). The issuer and the producer run their whole programs inside ONE elect block each (tcgen05 instructions issued from
// `if (lane == 0)` are wrapped in ELECT/branch loops by ptxas, and tcgen05.commit only tracks the MMAs of the committing
// thread); every hand-off is an mbarrier. Two input buffers and two halves of tensor memory: the tap MMAs of tile i+1 run
// while the workers drain tile i, and the TMA loads of tile i+2 are issued as soon as the MMAs of tile i have retired.
//   issuer, tile i:   taps(i) -> [stage 2 of tile i-1] -> extra(i) + commit D1(i) -> commit "buffer free"
//   workers, tile i:  [epilogue 2 of tile i-1] -> epilogue 1 of tile i          side channel, tile i: extras(i)
#ifdef LFSR_DEBUG_HOOKS
__device__ long long* g_em_dbg = nullptr;        // probe build: time stamps of the issuer at the start of its first 16 tiles
#define EM_STAMP(slot) do { if (g_em_dbg && blockIdx.x < 4096 && (slot) < 16) g_em_dbg[blockIdx.x * 16 + (slot)] = clock64(); } while (0)
// phases inside tile `EM_T` (second block of the buffer): issuer 0..4, epilogue worker 0 8..12, side channel 14, 15
#define EM_T 8
#define EM_PHASE(i, slot) do { if (g_em_dbg && (i) == EM_T && blockIdx.x < 4096) g_em_dbg[65536 + blockIdx.x * 16 + (slot)] = clock64(); } while (0)
#else
#define EM_STAMP(slot) do { } while (0)
#define EM_PHASE(i, slot) do { } while (0)
#endif
constexpr int kEmThreads = 448;
constexpr int kEmBufs = 3;            // input buffers: the side channel (and the TMA loads) run up to two tiles ahead of the MMAs
enum EmBar { EB_IMG = 0, EB_XF, EB_XE = EB_XF + kEmBufs, EB_AEX = EB_XE + kEmBufs, EB_D1 = EB_AEX + kEmBufs, EB_A2 = EB_D1 + 2, EB_D2 = EB_A2 + 2, EB_TE = EB_D2 + 2,
             EB_COUNT = EB_TE + 2 };

struct EmTile { int img, ty0, tx0; };
__device__ __forceinline__ EmTile em_decode(const EpiMmaArgs& a, int t) {
  EmTile e;
  e.tx0 = (t % a.tiles_x) * em::TW; t /= a.tiles_x;
  e.ty0 = (t % a.tiles_y) * a.R;
  e.img = t / a.tiles_y;
  return e;
}

__global__ void __launch_bounds__(kEmThreads, 1)
mel_epi_branch_mma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmE, const EpiMmaArgs a) {
  using namespace em;
  extern __shared__ uint8_t em_raw[];
  const uint32_t raw = smem_u32(em_raw);
  uint8_t* smem = em_raw + (((raw + 1023u) & ~1023u) - raw);
  const int ntap = a.ntap;
  const int img_bytes = image_bytes(ntap);
  uint8_t* Bt = smem;                                  // taps | extra | fuse | extra-channel tap weights (float2 per tap)
  uint8_t* Bex = Bt + ntap * kTapB;
  uint8_t* B2 = Bt + b2_offset(ntap);
  const float2* exw = reinterpret_cast<const float2*>(B2 + kB2);
  const int t16_bytes = (a.npx * 32 + 1023) & ~1023, t2_bytes = (a.npx * 16 + 1023) & ~1023;
  uint8_t* T16 = smem + ((img_bytes + 1023) & ~1023);  // 2 x [npx][16 fp16]  channels 0..15, SWIZZLE_32B rows
  uint8_t* T2h = T16 + kEmBufs * t16_bytes;            // x [npx][8 fp16]   channels 16..23 (16, 17 used)
  uint8_t* Aex = T2h + kEmBufs * t2_bytes;             // buffers x 2 blocks x [128][16 fp16]
  uint8_t* A2 = Aex + 2 * kEmBufs * kAex;                        // 2 blocks x [128][64 fp16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(A2 + 2 * kA2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + EB_COUNT);
  int* tapoff = reinterpret_cast<int*>(tmem_slot + 1);           // byte shift of the A view per tap
  uint64_t* adesc = reinterpret_cast<uint64_t*>(bars + 48);      // [2 buffers][2 blocks][kMaxTaps] A descriptors of the tap MMAs
  const int tid = threadIdx.x, warp = tid >> 5;
  const int P = a.P, halo = a.halo, half = a.KL / 2;
  const int total = a.total_tiles;
  const int n_my = ((int)total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) {
    mbar_init(bars + EB_IMG, 1);
    for (int i = 0; i < kEmBufs; ++i) {
      mbar_init(bars + EB_XF + i, 1); mbar_init(bars + EB_XE + i, 1);
      mbar_init(bars + EB_AEX + i, 128);      // one per input buffer: the side channel may finish tile i+1 before the issuer asks for tile i
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bars + EB_D1 + i, 1); mbar_init(bars + EB_A2 + i, 128); mbar_init(bars + EB_D2 + i, 1);
      mbar_init(bars + EB_TE + i, 256);
    }
    fence_barrier_init();
    mbar_expect_tx(bars + EB_IMG, (uint32_t)img_bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(Bt)), "l"(a.packed), "r"((uint32_t)img_bytes), "r"(smem_u32(bars + EB_IMG)) : "memory");
  }
  if (tid < ntap) {
    int dy = 0, dx = 0;
    if (tid < a.KL) dx = tid - half;
    else if (tid < 2 * a.KL) dy = tid - a.KL - half;
    else { const int k = tid - 2 * a.KL; dy = (k / 3 - 1) * a.dil; dx = (k % 3 - 1) * a.dil; }
    tapoff[tid] = (dy * P + dx) * 32;
    for (int bf = 0; bf < kEmBufs; ++bf)
      for (int m = 0; m < 2; ++m)
        adesc[(bf * 2 + m) * kMaxTaps + tid] =
            desc32(smem_u32(T16) + (uint32_t)(bf * t16_bytes) + (uint32_t)(a.G0 + m * 128) * 32u + (uint32_t)((dy * P + dx) * 32));
  }
  if (warp == 8) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 9) {
    // ================= TMA producer =================
    if (elect_one()) {
      const uint32_t tx = (uint32_t)((a.R + 2 * halo) * P * (32 + 16));
      for (int i = 0; i < n_my; ++i) {
        const int bf = i % kEmBufs;
        const EmTile e = em_decode(a, (int)blockIdx.x + i * (int)gridDim.x);
        mbar_wait(bars + EB_XE + bf, (((uint32_t)(i / kEmBufs)) & 1u) ^ 1u);  // the MMAs that read this buffer have retired
        mbar_expect_tx(bars + EB_XF + bf, tx);
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(smem_u32(T16 + bf * t16_bytes)), "l"(&tmX), "r"(smem_u32(bars + EB_XF + bf)), "r"(0), "r"(e.tx0 - halo),
                       "r"(e.ty0 - halo), "r"(e.img) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(smem_u32(T2h + bf * t2_bytes)), "l"(&tmE), "r"(smem_u32(bars + EB_XF + bf)), "r"(0), "r"(e.tx0 - halo),
                       "r"(e.ty0 - halo), "r"(e.img) : "memory");
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ================= MMA issuer =================
    if (elect_one()) {
      const uint32_t id1 = make_idesc(0, 32), idx = make_idesc(0, 96);
      const uint64_t dbt = desc32(smem_u32(Bt)), dbx = desc32(smem_u32(Bex)), db2 = make_smem_desc(smem_u32(B2));
      const int KL = a.KL;
      auto stage2 = [&](int i) {                     // fuse GEMM of tile i, block by block as its operand rows arrive
        const uint32_t th = tmem + (uint32_t)((i & 1) * 256);
#pragma unroll 1
        for (int m = 0; m < 2; ++m) {
          mbar_wait(bars + EB_A2 + m, (uint32_t)i & 1u);
          tc_fence_after();
          const uint64_t da = make_smem_desc(smem_u32(A2) + m * kA2);
          const uint32_t d = th + (uint32_t)(m * 96);
          umma_f16<0>(d, da, db2, id1);
          umma_f16<1>(d, da + 2, db2 + 2, id1);
          umma_f16<1>(d, da + 4, db2 + 4, id1);
          umma_f16<1>(d, da + 6, db2 + 6, id1);
          umma_commit(bars + EB_D2 + m);
        }
      };
      mbar_wait(bars + EB_IMG, 0);
      for (int i = 0; i < n_my; ++i) {
        const int bf = i % kEmBufs, hp = i & 1;
        const uint32_t th = tmem + (uint32_t)(hp * 256);
        EM_STAMP(i);
        mbar_wait(bars + EB_XF + bf, ((uint32_t)(i / kEmBufs)) & 1u);        // input tile landed
        mbar_wait(bars + EB_TE + hp, (((uint32_t)i >> 1) & 1u) ^ 1u);        // this half of tensor memory was drained (tile i-2)
        tc_fence_after();
        EM_PHASE(i, 0);
#pragma unroll 1
        for (int m = 0; m < 2; ++m) {
          const uint64_t* ad = adesc + (bf * 2 + m) * kMaxTaps;
          const uint32_t d0 = th + (uint32_t)(m * 96);
          int t = 0;
#pragma unroll 1
          for (int br = 0; br < 3; ++br) {
            const int n = br == 2 ? 9 : KL;
            const uint32_t d = d0 + (uint32_t)(br * 32);
            umma_f16<0>(d, ad[t], dbt + (uint64_t)(t * (kTapB >> 4)), id1);
            ++t;
#pragma unroll 4
            for (int k = 1; k < n; ++k, ++t) umma_f16<1>(d, ad[t], dbt + (uint64_t)(t * (kTapB >> 4)), id1);
          }
        }
        EM_PHASE(i, 1);
        if (i > 0) stage2(i - 1);
        EM_PHASE(i, 2);
        mbar_wait(bars + EB_AEX + bf, ((uint32_t)(i / kEmBufs)) & 1u);       // channels 16, 17 of this tile are in Aex
        EM_PHASE(i, 3);
        tc_fence_after();
#pragma unroll 1
        for (int m = 0; m < 2; ++m) {
          umma_f16<1>(th + (uint32_t)(m * 96), desc32(smem_u32(Aex) + (uint32_t)(bf * 2 + m) * kAex), dbx, idx);
          umma_commit(bars + EB_D1 + m);
        }
        umma_commit(bars + EB_XE + bf);                                      // input buffer free once all of this has retired
        EM_PHASE(i, 4);
      }
      if (n_my > 0) stage2(n_my - 1);
    }
    __syncwarp();
  } else if (warp < 8) {
    // ================= workers: the two epilogues =================
    const int m_blk = tid >> 7, r_blk = tid & 127;
    const int g = a.G0 + m_blk * 128 + r_blk;         // flattened padded position of this thread's pixel (= its MMA row)
    const int gy = g / P, gx = g - gy * P;
    const int ty = gy - halo, tx = gx - halo;
    const bool in_tile = tx >= 0 && tx < TW && ty >= 0 && ty < a.R;
    const uint32_t a2_row = smem_u32(A2) + (uint32_t)m_blk * kA2 + (uint32_t)r_blk * 128u;
    const uint32_t swz = (uint32_t)r_blk & 7u;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)m_blk * 96u;
    auto epilogue2 = [&](int i) {                      // LReLU(fuse) of tile i -> global
      const EmTile e = em_decode(a, (int)blockIdx.x + i * (int)gridDim.x);
      mbar_wait(bars + EB_D2 + m_blk, (uint32_t)i & 1u);
      tc_fence_after();
      float o[32];
      tmem_ld32(lane_base + (uint32_t)((i & 1) * 256), o);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(bars + EB_TE + (i & 1));             // this half of tensor memory may be overwritten
      const int oy = e.ty0 + ty, ox = e.tx0 + tx;
      if (in_tile && ox < a.in.w && oy < a.in.h) {
        float* dst = a.out.p + a.out.pix(e.img, oy, ox);
#pragma unroll
        for (int k = 0; k < EC / 2; ++k) {
          float x0 = o[2 * k], x1 = o[2 * k + 1];
          x0 = x0 > 0.f ? x0 : x0 * a.slope;
          x1 = x1 > 0.f ? x1 : x1 * a.slope;
          reinterpret_cast<float2*>(dst)[k] = make_float2(x0, x1);
        }
      }
    };
    for (int i = 0; i < n_my; ++i) {
      const int bf = i & 1;                            // half of tensor memory
      if (tid == 0) EM_PHASE(i, 8);
      if (i > 0) epilogue2(i - 1);
      if (tid == 0) EM_PHASE(i, 9);
      // ---- stage 2 operand: LReLU(stage 1) as fp16
      mbar_wait(bars + EB_D1 + m_blk, (uint32_t)i & 1u);
      if (tid == 0) EM_PHASE(i, 10);
      tc_fence_after();
      {
        const uint32_t tlane = lane_base + (uint32_t)(bf * 256);
        uint32_t pk[32];
#pragma unroll
        for (int k = 27; k < 32; ++k) pk[k] = 0u;
#pragma unroll
        for (int br = 0; br < 3; ++br) {
          float s[32];
          tmem_ld32(tlane + br * 32, s);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            float x0 = s[2 * k], x1 = s[2 * k + 1];
            x0 = x0 > 0.f ? x0 : x0 * a.slope;
            x1 = x1 > 0.f ? x1 : x1 * a.slope;
            pk[br * 9 + k] = pack_f16x2(x0, x1);
          }
        }
        et::store_row(a2_row, swz, pk);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bars + EB_A2 + m_blk);
      if (tid == 0) EM_PHASE(i, 11);
    }
    if (n_my > 0) epilogue2(n_my - 1);
  } else {
    // ================= side channel (4 warps): channels 16, 17 on the CUDA cores, one tile ahead of the epilogues =================
    const int r_e = tid - 320;                         // 0..127: row r_e of both 128-row blocks
    mbar_wait(bars + EB_IMG, 0);                       // the extra-channel tap weights live in the operand image
    for (int i = 0; i < n_my; ++i) {
      const int bf = i % kEmBufs;
      mbar_wait(bars + EB_XF + bf, ((uint32_t)(i / kEmBufs)) & 1u);
      if (tid == 320) EM_PHASE(i, 14);
      // (Aex[bf] was last read by the extra MMA of tile i-3, which retired before this tile's input buffer was refilled)
#pragma unroll 1
      for (int m = 0; m < 2; ++m) {
        const int g = a.G0 + m * 128 + r_e;
        const __half2* tp = reinterpret_cast<const __half2*>(T2h + bf * t2_bytes) + (size_t)g * 4;     // 8 halves per pixel
        f32x2 acc[3];
        int t = 0;
#pragma unroll
        for (int br = 0; br < 3; ++br) {
          const int n = br == 2 ? 9 : a.KL;
          f32x2 s0 = pack2(0.f, 0.f), s1 = s0;
          int k = 0;
#pragma unroll 4
          for (; k + 1 < n; k += 2, t += 2) {
            const float2 v0 = __half22float2(tp[(tapoff[t] >> 5) * 4]), v1 = __half22float2(tp[(tapoff[t + 1] >> 5) * 4]);
            s0 = fma2(pack2(v0.x, v0.y), *reinterpret_cast<const f32x2*>(exw + t), s0);
            s1 = fma2(pack2(v1.x, v1.y), *reinterpret_cast<const f32x2*>(exw + t + 1), s1);
          }
          if (k < n) {
            const float2 v0 = __half22float2(tp[(tapoff[t] >> 5) * 4]);
            s0 = fma2(pack2(v0.x, v0.y), *reinterpret_cast<const f32x2*>(exw + t), s0);
            ++t;
          }
          float a0, a1, b0, b1;
          unpack2(s0, a0, a1); unpack2(s1, b0, b1);
          acc[br] = pack2(a0 + b0, a1 + b1);
        }
        float x0, x1, x2, x3, x4, x5;
        unpack2(acc[0], x0, x1); unpack2(acc[1], x2, x3); unpack2(acc[2], x4, x5);
        const uint32_t aex_row = smem_u32(Aex) + (uint32_t)(bf * 2 + m) * kAex + (uint32_t)r_e * 32u;
        st_shared_v4(unit32(aex_row, 0), pack_f16x2(x0, x1), pack_f16x2(x2, x3), pack_f16x2(x4, x5), 0u);
        st_shared_v4(unit32(aex_row, 1), 0u, 0u, 0u, 0u);
      }
      fence_proxy_async();
      if (tid == 320) EM_PHASE(i, 15);
      mbar_arrive(bars + EB_AEX + bf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, 512);
}

}  // namespace lfsr

using namespace lfsr;

static int small_cs(int c) { int cs = c; while (cs % 8 != 4) ++cs; return cs; }
static size_t small_smem(const lfsr_tensor* in, const lfsr_conv_desc* d, int cout) {
  const int cs = small_cs(in->c);
  const int hw = ST_W + d->dil_w * (d->kw - 1), hh = ST_H + d->dil_h * (d->kh - 1);
  return ((size_t)hh * hw * cs + (size_t)d->kh * d->kw * cout * cs) * sizeof(float);
}

extern "C" int lfsr_conv2d_small_cout_supported(const lfsr_tensor* in, const lfsr_tensor* out, const lfsr_conv_desc* d) {
  if (!tensor_ok(in) || !tensor_ok(out) || !d) return 0;
  if (out->c > 4 || d->stride_h != 1 || d->stride_w != 1 || d->in_perm || d->out_perm || d->mul.ptr || d->in_scale) return 0;
  if (d->shuf_ry > 1 || d->shuf_rx > 1 || d->block_h > 0 || d->block_w > 0) return 0;
  if (2 * d->pad_h != d->dil_h * (d->kh - 1) || 2 * d->pad_w != d->dil_w * (d->kw - 1)) return 0;
  if (out->n != in->n || out->h != in->h || out->w != in->w) return 0;
  if (in->c < 8 || (in->c & 1) || (in->ld & 1) || ((uintptr_t)in->ptr & 7)) return 0;
  if (small_smem(in, d, out->c) > 110 * 1024) return 0;
  return 1;
}

extern "C" int lfsr_conv2d_small_cout(const lfsr_tensor* in, const float* w_packed, const lfsr_tensor* out,
                                      const lfsr_conv_desc* d, void* stream) {
  LFSR_REQUIRE(w_packed && lfsr_conv2d_small_cout_supported(in, out, d), "lfsr_conv2d_small_cout: unsupported problem");
  SmallArgs a;
  a.in = view_of(in); a.out = view_of(out);
  a.res = d->res.ptr ? view_of(&d->res) : null_view();
  if (d->res.ptr)
    LFSR_REQUIRE(d->res.n == out->n && d->res.h == out->h && d->res.w == out->w && d->res.c == out->c,
                 "lfsr_conv2d_small_cout: res geometry");
  a.w = w_packed; a.bias = d->bias;
  a.kh = d->kh; a.kw = d->kw; a.dh = d->dil_h; a.dw = d->dil_w; a.ph = d->pad_h; a.pw = d->pad_w;
  a.cout = out->c; a.act = d->act; a.slope = d->act_slope; a.alpha = d->alpha;
  a.tiles_x = ceil_div(out->w, ST_W); a.tiles_y = ceil_div(out->h, ST_H);
  a.vec16 = (in->c % 4 == 0) && (in->ld % 4 == 0) && (((uintptr_t)in->ptr & 15) == 0);
  const int blocks = out->n * a.tiles_x * a.tiles_y;
  const int cs = small_cs(in->c);
  const size_t smem = small_smem(in, d, out->c);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_SMALL(CO)                                                                                         \
  do {                                                                                                           \
    static DevOnce once;                                                                                         \
    if (once.need()) {                                                                                           \
      if (opt_in_smem(conv_small_cout_kernel<CO>, 110 * 1024, "lfsr_conv2d_small_cout")) return LFSR_ERR_CUDA;   \
      once.done();                                                                                               \
    }                                                                                                            \
    conv_small_cout_kernel<CO><<<blocks, 256, smem, st>>>(a, cs);                                                \
  } while (0)
  switch (out->c) {
    case 1: LAUNCH_SMALL(1); break;
    case 2: LAUNCH_SMALL(2); break;
    case 3: LAUNCH_SMALL(3); break;
    default: LAUNCH_SMALL(4); break;
  }
#undef LAUNCH_SMALL
  return check_launch("conv_small_cout_kernel");
}

extern "C" int lfsr_mel_epi_branch(const lfsr_tensor* in, const float* w_packed, const lfsr_tensor* out, int klen, int dil,
                                   float slope, void* stream) {
  LFSR_REQUIRE(tensor_ok(in) && tensor_ok(out) && w_packed, "lfsr_mel_epi_branch: null/invalid tensor");
  LFSR_REQUIRE(in->c == EC && out->c == EC, "lfsr_mel_epi_branch: built for %d-channel EPI splits, got %d", EC, in->c);
  LFSR_REQUIRE(in->n == out->n && in->h == out->h && in->w == out->w, "lfsr_mel_epi_branch: shape mismatch");
  LFSR_REQUIRE(klen > 0 && (klen & 1) && klen <= 31 && dil > 0, "lfsr_mel_epi_branch: bad kernel length");
  LFSR_REQUIRE(in->ld % 4 == 0 && in->ld >= EC + 2 && ((uintptr_t)in->ptr & 15) == 0,
               "lfsr_mel_epi_branch: input slice must be 16-byte aligned with 2 readable pad floats (grouped trunk layout)");
  LFSR_REQUIRE(out->ld % 2 == 0 && ((uintptr_t)out->ptr & 7) == 0, "lfsr_mel_epi_branch: 8-byte aligned output slice required");
  EpiArgs a;
  a.in = view_of(in); a.out = view_of(out); a.w = w_packed; a.KL = klen; a.dil = dil; a.slope = slope;
  a.tiles_x = ceil_div(in->w, EP_W); a.tiles_y = ceil_div(in->h, EP_H);
  const int halo = klen / 2 > dil ? klen / 2 : dil;
  const size_t smem = ((size_t)(2 * klen + 9 + 6 * EC) * EP_CS + (size_t)(EP_W + 2 * halo) * (EP_H + 2 * halo) * EP_CS) * sizeof(float);
  LFSR_REQUIRE(smem <= 200 * 1024, "lfsr_mel_epi_branch: kernel length / dilation too large for the staged tile");
  static DevOnce once;
  if (once.need()) {
    if (opt_in_smem(mel_epi_branch_kernel, 200 * 1024, "lfsr_mel_epi_branch")) return LFSR_ERR_CUDA;
    once.done();
  }
  mel_epi_branch_kernel<<<in->n * a.tiles_x * a.tiles_y, 256, smem, (cudaStream_t)stream>>>(a);
  return check_launch("mel_epi_branch_kernel");
}

static size_t epi_tc_smem(int klen, int dil) {
  const int halo = klen / 2 > dil ? klen / 2 : dil;
  size_t tile = (size_t)(EP_W + 2 * halo) * (EP_H + 2 * halo) * EP_CS * sizeof(float);
  if (tile < 2 * et::kA) tile = 2 * et::kA;
  const size_t taps = (size_t)(((2 * klen + 9) * EP_CS + 3) & ~3) * sizeof(float);
  return 1024 + et::kB1 + et::kB2 + tile + taps + 32;
}

extern "C" int lfsr_mel_epi_branch_tc(const lfsr_tensor* in, const float* w_packed, const lfsr_tensor* out, int klen, int dil,
                                      float slope, void* stream) {
  LFSR_REQUIRE(tensor_ok(in) && tensor_ok(out) && w_packed, "lfsr_mel_epi_branch_tc: null/invalid tensor");
  LFSR_REQUIRE(in->c == EC && out->c == EC, "lfsr_mel_epi_branch_tc: built for %d-channel EPI splits, got %d", EC, in->c);
  LFSR_REQUIRE(in->n == out->n && in->h == out->h && in->w == out->w, "lfsr_mel_epi_branch_tc: shape mismatch");
  LFSR_REQUIRE(klen > 0 && (klen & 1) && klen <= 31 && dil > 0, "lfsr_mel_epi_branch_tc: bad kernel length");
  LFSR_REQUIRE(in->ld % 4 == 0 && in->ld >= EC + 2 && ((uintptr_t)in->ptr & 15) == 0,
               "lfsr_mel_epi_branch_tc: input slice must be 16-byte aligned with 2 readable pad floats (grouped trunk layout)");
  LFSR_REQUIRE(out->ld % 2 == 0 && ((uintptr_t)out->ptr & 7) == 0, "lfsr_mel_epi_branch_tc: 8-byte aligned output slice required");
  EpiArgs a;
  a.in = view_of(in); a.out = view_of(out); a.w = w_packed; a.KL = klen; a.dil = dil; a.slope = slope;
  a.tiles_x = ceil_div(in->w, EP_W); a.tiles_y = ceil_div(in->h, EP_H);
  const size_t smem = epi_tc_smem(klen, dil);
  LFSR_REQUIRE(smem <= 200 * 1024, "lfsr_mel_epi_branch_tc: kernel length / dilation too large for the staged tile");
  static DevOnce once;
  if (once.need()) {
    if (opt_in_smem(mel_epi_branch_tc_kernel, 200 * 1024, "lfsr_mel_epi_branch_tc")) return LFSR_ERR_CUDA;
    once.done();
  }
  mel_epi_branch_tc_kernel<<<in->n * a.tiles_x * a.tiles_y, 256, smem, (cudaStream_t)stream>>>(a);
  return check_launch("mel_epi_branch_tc_kernel");
}

// ---- operand image of the all-tensor-core EPI kernel (host side) ----------------------------------------------------
static uint16_t f32_to_f16_bits(float f) {       // round to nearest even, saturating to the finite range, denormals kept
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return (uint16_t)(sign | (x > 0x7f800000u ? 0x7e00u : 0x7bffu));
  if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu);
  if (x < 0x38800000u) {                          // subnormal half
    if (x < 0x33000000u) return (uint16_t)sign;
    const int e = (int)(x >> 23);
    uint32_t m = (x & 0x7fffffu) | 0x800000u;
    const int shift = 126 - e;                    // 14 .. 24
    const uint32_t r = m >> shift, rem = m & ((1u << shift) - 1), halfway = 1u << (shift - 1);
    return (uint16_t)(sign | (r + ((rem > halfway || (rem == halfway && (r & 1))) ? 1 : 0)));
  }
  const uint32_t r = x - 0x38000000u;
  const uint32_t rem = r & 0x1fffu;
  uint32_t h = r >> 13;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;
  return (uint16_t)(sign | h);
}

extern "C" size_t lfsr_mel_epi_pack_bytes(int klen) {
  if (klen <= 0 || !(klen & 1) || 2 * klen + 9 > em::kMaxTaps) return 0;
  return (size_t)em::image_bytes(2 * klen + 9);
}

extern "C" int lfsr_mel_epi_pack(const float* w, void* out, int klen) {
  LFSR_REQUIRE(w && out && lfsr_mel_epi_pack_bytes(klen) > 0, "lfsr_mel_epi_pack: bad arguments (klen must be odd, 2 klen + 9 <= %d)", em::kMaxTaps);
  const int ntap = 2 * klen + 9;
  uint8_t* img = (uint8_t*)out;
  memset(img, 0, lfsr_mel_epi_pack_bytes(klen));
  const float* dw = w;                              // [ntap][EC]: dw_h | dw_v | dw_d
  const float* pw = w + ntap * EC;                  // 3 x [EC in][EC out]
  const float* fu = pw + 3 * EC * EC;               // [3*EC in][EC out]
  auto put32 = [](uint8_t* blk, int n, int k, float v) {       // [rows][16] fp16, SWIZZLE_32B (block 256-byte aligned)
    const uint32_t row = (uint32_t)n * 32u;
    const uint32_t off = row + ((((uint32_t)k >> 3) ^ ((row >> 7) & 1u)) << 4) + ((uint32_t)k & 7u) * 2u;
    const uint16_t h = f32_to_f16_bits(v);
    memcpy(blk + off, &h, 2);
  };
  for (int t = 0; t < ntap; ++t) {
    const int br = t < klen ? 0 : (t < 2 * klen ? 1 : 2);
    uint8_t* blk = img + (size_t)t * em::kTapB;
    for (int o = 0; o < EC; ++o)
      for (int c = 0; c < 16; ++c) put32(blk, o, c, dw[t * EC + c] * pw[br * EC * EC + c * EC + o]);
  }
  uint8_t* bex = img + (size_t)ntap * em::kTapB;
  for (int br = 0; br < 3; ++br)
    for (int o = 0; o < EC; ++o)
      for (int j = 0; j < 2; ++j) put32(bex, 32 * br + o, 2 * br + j, pw[br * EC * EC + (16 + j) * EC + o]);
  uint8_t* b2 = img + em::b2_offset(ntap);
  for (int n = 0; n < EC; ++n)
    for (int k = 0; k < 3 * EC; ++k) {                // [32 out][64 in] fp16, SWIZZLE_128B
      const uint32_t off = (uint32_t)n * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
      const uint16_t h = f32_to_f16_bits(fu[k * EC + n]);
      memcpy(b2 + off, &h, 2);
    }
  float* exw = (float*)(b2 + em::kB2);
  for (int t = 0; t < ntap; ++t) { exw[2 * t] = dw[t * EC + 16]; exw[2 * t + 1] = dw[t * EC + 17]; }
  return LFSR_OK;
}

typedef CUresult (*EmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EmEncodeFn em_get_encode() {
  static EmEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EmEncodeFn>(p);
  });
  return fn;
}

extern "C" int lfsr_mel_epi_branch_mma(const lfsr_tensor* in, const lfsr_tensor* in16, const void* packed, const lfsr_tensor* out,
                                       int klen, int dil, float slope, void* stream) {
  LFSR_REQUIRE(tensor_ok(in) && tensor_ok(out) && tensor_ok(in16) && packed, "lfsr_mel_epi_branch_mma: null/invalid tensor");
  LFSR_REQUIRE(in16->n == in->n && in16->h == in->h && in16->w == in->w && in16->c >= EC && in16->ld % 8 == 0 && in16->ld >= 24 &&
                   ((uintptr_t)in16->ptr & 15) == 0,
               "lfsr_mel_epi_branch_mma: in16 must be the fp16 copy of `in` (18 channels + readable pad up to 24, 16-byte aligned pixels)");
  LFSR_REQUIRE(in->c == EC && out->c == EC, "lfsr_mel_epi_branch_mma: built for %d-channel EPI splits, got %d", EC, in->c);
  LFSR_REQUIRE(in->n == out->n && in->h == out->h && in->w == out->w, "lfsr_mel_epi_branch_mma: shape mismatch");
  LFSR_REQUIRE(lfsr_mel_epi_pack_bytes(klen) > 0 && dil > 0, "lfsr_mel_epi_branch_mma: bad kernel length");
  LFSR_REQUIRE(in->ld % 4 == 0 && in->ld >= EC + 2 && ((uintptr_t)in->ptr & 15) == 0 && ((uintptr_t)packed & 15) == 0,
               "lfsr_mel_epi_branch_mma: input slice must be 16-byte aligned with 2 readable pad floats (grouped trunk layout)");
  LFSR_REQUIRE(out->ld % 2 == 0 && ((uintptr_t)out->ptr & 7) == 0, "lfsr_mel_epi_branch_mma: 8-byte aligned output slice required");
  EpiMmaArgs a;
  a.in = view_of(in); a.out = view_of(out); a.packed = (const uint8_t*)packed; a.KL = klen; a.dil = dil; a.slope = slope;
  a.halo = klen / 2 > dil ? klen / 2 : dil;
  a.ntap = 2 * klen + 9;
  a.P = em::TW + 2 * a.halo;
  a.G0 = a.halo * a.P + a.halo;
  a.R = (256 - em::TW) / a.P + 1;                    // output rows whose 32 pixels all lie inside the 256 MMA rows
  LFSR_REQUIRE(a.R >= 1, "lfsr_mel_epi_branch_mma: halo too large");
  a.npx = (2 * a.G0 + 256 + 31) & ~31;               // last MMA row + largest shift, rounded up
  a.tiles_x = ceil_div(in->w, em::TW); a.tiles_y = ceil_div(in->h, a.R);
  const long long total = (long long)in->n * a.tiles_x * a.tiles_y;
  LFSR_REQUIRE(total <= 0x7fffffffLL, "lfsr_mel_epi_branch_mma: too many tiles");
  a.total_tiles = (int)total;
  const size_t img_b = ((size_t)em::image_bytes(a.ntap) + 1023) & ~(size_t)1023;
  const size_t t16_b = ((size_t)a.npx * 32 + 1023) & ~(size_t)1023, t2_b = ((size_t)a.npx * 16 + 1023) & ~(size_t)1023;
  const size_t smem = 1024 + img_b + kEmBufs * (t16_b + t2_b + 2 * em::kAex) + 2 * em::kA2 + 384 + 2 * kEmBufs * em::kMaxTaps * 8 + 64;
  LFSR_REQUIRE(smem <= 225 * 1024, "lfsr_mel_epi_branch_mma: kernel length / dilation too large for the staged tile");
  static DevOnce once;
  if (once.need()) {
    if (opt_in_smem(mel_epi_branch_mma_kernel, 225 * 1024, "lfsr_mel_epi_branch_mma")) return LFSR_ERR_CUDA;
    once.done();
  }
  EmEncodeFn encode = em_get_encode();
  LFSR_REQUIRE(encode, "lfsr_mel_epi_branch_mma: cuTensorMapEncodeTiled is not available");
  CUtensorMap tmX;
  {
    const cuuint64_t ld_b = (cuuint64_t)in16->ld * 2;
    cuuint64_t dims[4] = {16, (cuuint64_t)in->w, (cuuint64_t)in->h, (cuuint64_t)in->n};
    cuuint64_t strides[3] = {ld_b, ld_b * in->w, ld_b * in->w * in->h};
    cuuint32_t box[4] = {16, (cuuint32_t)a.P, (cuuint32_t)(a.R + 2 * a.halo), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    LFSR_REQUIRE(box[1] <= 256 && box[2] <= 256, "lfsr_mel_epi_branch_mma: tile too large for one tensor load");
    CUresult r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, in16->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("lfsr_mel_epi_branch_mma: cuTensorMapEncodeTiled failed with %d", (int)r); return LFSR_ERR_CUDA; }
  }
  CUtensorMap tmE;
  {
    const cuuint64_t ld_b = (cuuint64_t)in16->ld * 2;
    cuuint64_t dims[4] = {8, (cuuint64_t)in->w, (cuuint64_t)in->h, (cuuint64_t)in->n};
    cuuint64_t strides[3] = {ld_b, ld_b * in->w, ld_b * in->w * in->h};
    cuuint32_t box[4] = {8, (cuuint32_t)a.P, (cuuint32_t)(a.R + 2 * a.halo), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&tmE, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, (char*)in16->ptr + 32, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("lfsr_mel_epi_branch_mma: cuTensorMapEncodeTiled(extra channels) failed with %d", (int)r); return LFSR_ERR_CUDA; }
  }
  const int nsm = sm_count_current();
  const int grid = a.total_tiles < nsm ? a.total_tiles : nsm;
  mel_epi_branch_mma_kernel<<<grid, kEmThreads, smem, (cudaStream_t)stream>>>(tmX, tmE, a);
  return check_launch("mel_epi_branch_mma_kernel");
}

#ifdef LFSR_DEBUG_HOOKS
extern "C" int lfsr_debug_set_em_dbg(long long* dev_ptr) {
  return cudaMemcpyToSymbol(lfsr::g_em_dbg, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? LFSR_OK : LFSR_ERR_CUDA;
}
#endif
