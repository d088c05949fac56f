// EPIT's BasicTrans (EPIT.py:74-128, mask :93-108) as ONE tcgen05/TMEM kernel.
//
//   x [T, 64] --linear_in--> X [128] --LN--> N --(Wq, Wk)--> Q, K ;  V = X Wv^T  (v comes from the un-normed tokens, :117-122)
//   per head (8 x 16): softmax(Q K^T / 4 + band mask) V ;  X2 = attn Wo^T + X ;  X3 = W2 relu(W1 LN(X2)) + X2 ;  y = X3 Wout^T
//
// Work item = HALF an EPI sequence. A sequence is L = A*S tokens (A angular x S spatial positions of one EPI line); the
// additive mask of the reference (mask_field [2A, 11]) lets token (a, s) attend all A rows of the positions |s - s'| <= 5.
// With the tokens ordered s-major (row = s*A + a) a tile that owns the query positions [qs0, qs1) needs only the
// positions [qs0 - 5, qs1 + 5): for S = 32, A = 5 two tiles of 16 query positions load 21 positions = 105 token rows
// each, which fits the 128 TMEM lanes / 128 MMA rows, so queries, keys and values of a tile are the same 105 rows and
// nothing ever leaves the SM between linear_in and linear_out (HBM: 256 B in + 256 B out per token instead of ~13 KB
// for the ten separate launches this replaces).
//
// GEMM plan per tile (M = 128 rows, accumulators in TMEM, fp32):
//   R  [cols   0..127]  residual stream: X = x W_in^T (kind::tf32, A tile straight from TMA), then += O Wo^T, += F W2^T
//   T1 [cols 128..255]  V^T = Wv X^T (weights as the A operand -> lanes = channels, so the epilogue writes K-major V^T rows
//                       without a transpose), later O (8 heads x 16 columns), later y (64 columns)
//   T2 [cols 256..511]  [Q | K], later the two ping-pong score tiles S (N = keys), later relu-input F (256 columns)
// Every operand produced inside the kernel is written by the epilogue warps as fp16 (10-bit mantissa = TF32's precision,
// fp32 accumulate) in the K-major SWIZZLE_128B layout the UMMA descriptors read; kind::f16 halves shared-memory bytes
// and MMA time against tf32, which is what lets a whole tile live in 160 KB. The 304 KB of weights per tile stream
// through a 4-slot TMA ring from L2.
// Warps: 0..7 epilogue (two per TMEM lane quarter: they split columns / heads), 8 TMA producer, 9 MMA issuer.
#include <cuda.h>
#include <cuda_fp16.h>
#include <string.h>
#include <mutex>
#include "lfsr_common.cuh"
#include "lfsr_ptx.cuh"

namespace lfsr {
namespace bt {
using namespace lfsr::ptx;

constexpr int kEpiWarps = 16, kTmaWarp = 16, kMmaWarp = 17, kThreads = 32 * 18;
constexpr int kE = 128, kC = 64, kHeads = 8, kHd = 16;
constexpr int kRows = 112;                  // token rows a tile may hold (7 groups of 16 score columns)
constexpr int kMaxKeys = kRows;
constexpr int kGrpPerWarp = 3;              // 16-column score groups a softmax warp keeps in registers (host checks the plan)
constexpr int kChunk = kRows * 128;         // one K-chunk (64 fp16 / 32 tf32 per row) of a token-row operand: 14 KB
constexpr int kRegion = 2 * kChunk;         // 28 KB
constexpr int kVtChunk = 128 * 128, kVtRegion = 2 * kVtChunk;      // V^T: rows are the 128 channels
constexpr int kSlot = 16384, kSlots = 4;
constexpr int kMaskBytes = kRows * 64;      // band-mask operands: [112 rows][32 fp16], SWIZZLE_64B K-major, one for A, one for B
                                            // (the M = 128 MMA reads 16 more A rows: whatever follows, they only reach unused lanes)
constexpr float kMaskNeg = -30000.f;        // added (log2 domain) to the scores of keys outside the band: exp2 -> 0
constexpr int kBlocks = 18;                 // weight blocks (16 KB each) per tile, in consumption order
constexpr int kXchgFloats = 8 * kRows;      // LN partials [4 parts][112 rows] float2 / softmax partial max, sum [2 pairs][2][112] each
constexpr int kSmemBytes = 4 * kRegion + kVtRegion + kSlots * kSlot + 2 * kMaskBytes + kXchgFloats * 4 + 40 * 8 /* barriers */ + 16;
// weight block indices
constexpr int kWin = 0, kWv = 1, kWq = 3, kWk = 5, kWo = 7, kW1 = 9, kW2 = 13, kWout = 17;

struct Params {
  int A, S, w, SL, nrows, nkeys, nt, sq;
  int nseq, npq, nq, total_tiles;
  long long stride_a, stride_s, stride_b, stride_p, stride_q;      // in tokens (pixels)
  float* y;
  int ld_y;
  float eps1, eps2, qscale;
  long long* dbg;                        // probe build only: per-CTA cycle accounting (profiles/probe_bt_phases.py)
  float ln[4][kE];                       // gamma1, beta1, gamma2, beta2 (constant bank: uniform loads in the LN loops)
};

#ifdef LFSR_DEBUG_HOOKS
#define BT_MWAIT(i, bar, par_) do { const long long t0_ = clock64(); mbar_wait(bar, par_); if (p.dbg) p.dbg[blockIdx.x * 64 + 32 + (i)] += clock64() - t0_; } while (0)
#else
#define BT_MWAIT(i, bar, par_) mbar_wait(bar, par_)
#endif

enum Bar {
  X_FULL = 0, X_EMPTY, W_FULL, W_EMPTY = W_FULL + kSlots, R_FULL = W_EMPTY + kSlots, VT_FULL, K_FULL, S_FULL,
  P_EMPTY = S_FULL + 3, F_FULL = P_EMPTY + 2, Y_FULL = F_FULL + 2,
  // arrivals of the epilogue warps
  X_READY, N_READY, QKV_READY, P_FULL, O_READY = P_FULL + 2, N2_READY, F_READY, X3_READY = F_READY + 2, NUM_BARS
};
static_assert(NUM_BARS <= 40, "barrier block");

// `n` (multiple of 8) floats -> fp16 into 16-byte units u0, u0+1, .. of one 128-byte K-major SWIZZLE_128B row; the chunk
// starts at a 1024-byte boundary
template <int N>
__device__ __forceinline__ void store_f16(uint32_t chunk, int row, int u0, const float* v) {
  const uint32_t rb = chunk + (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll
  for (int j = 0; j < N / 8; ++j)
    st_shared_v4(rb + ((((uint32_t)(u0 + j)) ^ sw) << 4), pack_f16x2(v[8 * j], v[8 * j + 1]), pack_f16x2(v[8 * j + 2], v[8 * j + 3]),
                 pack_f16x2(v[8 * j + 4], v[8 * j + 5]), pack_f16x2(v[8 * j + 6], v[8 * j + 7]));
}
template <int N>
__device__ __forceinline__ void store_zero(uint32_t chunk, int row, int u0) {
  const uint32_t rb = chunk + (uint32_t)row * 128u;
  const uint32_t sw = (uint32_t)(row & 7);
#pragma unroll
  for (int j = 0; j < N / 8; ++j) st_shared_v4(rb + ((((uint32_t)(u0 + j)) ^ sw) << 4), 0u, 0u, 0u, 0u);
}

// four K-steps (one 128-byte chunk row) of a kind::f16 / kind::tf32 GEMM slice
template <bool TF32, bool FIRST>
__device__ __forceinline__ void mma_chunk(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
  if (TF32) {
    umma_tf32<FIRST ? 0 : 1>(d, a, b, idesc);
    umma_tf32<1>(d, a + 2, b + 2, idesc); umma_tf32<1>(d, a + 4, b + 4, idesc); umma_tf32<1>(d, a + 6, b + 6, idesc);
  } else {
    umma_f16<FIRST ? 0 : 1>(d, a, b, idesc);
    umma_f16<1>(d, a + 2, b + 2, idesc); umma_f16<1>(d, a + 4, b + 4, idesc); umma_f16<1>(d, a + 6, b + 6, idesc);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
basictrans_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmY, const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* RA = smem;                 // x (fp32, TMA) / X fp16 / P of odd heads
  uint8_t* RB = smem + kRegion;       // N / P of even heads / N2 / X3
  uint8_t* RC = smem + 2 * kRegion;   // Q / O (head by head, each O_h into the slot of the dead Q_h)
  uint8_t* RD = smem + 3 * kRegion;   // K / F chunks 0, 1
  uint8_t* RE = smem + 4 * kRegion;   // V^T (128 channel rows) / F chunks 2, 3
  uint8_t* ring = RE + kVtRegion;
  uint8_t* MA = ring + kSlots * kSlot;          // one-hot(position of the query row)            [128][32] fp16
  uint8_t* MB = MA + kMaskBytes;                 // 0 / kMaskNeg per (key row, query position)     [128][32] fp16
  float* xchg = reinterpret_cast<float*>(MB + kMaskBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(xchg + kXchgFloats);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NUM_BARS; ++i) {
      uint32_t cnt = 1;
      if (i == X_READY || i == N_READY || i == QKV_READY || i == O_READY || i == N2_READY || i == X3_READY) cnt = kEpiWarps;
      if (i == P_FULL || i == P_FULL + 1 || i == F_READY || i == F_READY + 1) cnt = kEpiWarps / 2;
      mbar_init(bars + i, cnt);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmY);
  }
  // The band mask rides on the tensor pipe: S_h = Q_h K_h^T + onehot(s_q) . M^T, where M[key][s] = 0 if the key's position is
  // within half_window of s (and the key row exists), else kMaskNeg. Two extra K = 16 MMAs per head replace a compare +
  // select per score element on the CUDA cores. Operands are constants of the launch: built once per CTA.
  for (int i = threadIdx.x; i < kRows * 32; i += kThreads) {
    const int row = i >> 5, k = i & 31;
    const int sl = row / p.A;
    const bool is_tok = row < p.nrows;
    const float a = (is_tok && k == sl) ? 1.f : 0.f;
    const int dk = sl - k;
    const float b = (is_tok && k < p.SL && dk <= p.w && -dk <= p.w) ? 0.f : kMaskNeg;
    const uint32_t off = (uint32_t)(row >> 3) * 512u + (uint32_t)(row & 7) * 64u +
                         ((((uint32_t)k >> 3) ^ (((uint32_t)row >> 1) & 3u)) << 4) + ((uint32_t)k & 7u) * 2u;
    *reinterpret_cast<__half*>(MA + off) = __float2half_rn(a);
    *reinterpret_cast<__half*>(MB + off) = __float2half_rn(b);
  }
  fence_proxy_async();
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: R 0..127 | T1 128..255 (V^T, then score buffer 2, then y) | T2 256..511 ([Q|K], then score buffers 0, 1 at
  // +0 / +112 and the two 16-column O buffers at +224 / +240, then the 256 FFN columns)
  const uint32_t tR = tmem, tT1 = tmem + 128, tT2 = tmem + 256;
  const int n_my = ((int)p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

  if (warp == kTmaWarp) {
    // ================= TMA producer: x tiles + the weight ring =================
    if (elect_one()) {
      uint32_t wcnt = 0;
      const uint32_t x_bytes = (uint32_t)p.nrows * 128u;              // [rows][64 fp16]
      auto load_x = [&](int tile) {
        const int seq = tile / p.nt, t = tile - seq * p.nt;
        const int b = seq / p.npq, pq = seq - b * p.npq;
        int ls0 = t * p.sq - p.w;
        if (ls0 < 0) ls0 = 0;
        if (ls0 > p.S - p.SL) ls0 = p.S - p.SL;
        mbar_expect_tx(bars + X_FULL, x_bytes);
        tma_load_5d(RA, &tmX, bars + X_FULL, 0, 0, ls0, pq, b);
      };
      auto load_w = [&](int blk) {
        const uint32_t slot = wcnt % kSlots, use = wcnt / kSlots;
        mbar_wait(bars + W_EMPTY + slot, (use & 1u) ^ 1u);
        mbar_expect_tx(bars + W_FULL + slot, kSlot);
        tma_load_2d(ring + slot * kSlot, &tmW, bars + W_FULL + slot, 0, blk * 128);
        ++wcnt;
      };
      if (n_my > 0) load_x((int)blockIdx.x);
      for (int it = 0; it < n_my; ++it) {
        for (int blk = 0; blk < 11; ++blk) load_w(blk);
        if (it + 1 < n_my) {              // region RA is free once the last P.V MMA of this tile has retired
          mbar_wait(bars + X_EMPTY, (uint32_t)it & 1u);
          load_x((int)blockIdx.x + (it + 1) * (int)gridDim.x);
        }
        for (int blk = 11; blk < kBlocks; ++blk) load_w(blk);
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer: one elected lane walks the whole tile program =================
    if (elect_one()) {
      uint32_t wcnt = 0;
      const uint32_t id_f16 = make_idesc(0, 128), id_s = make_idesc(0, p.nkeys),
                     id_pv = make_idesc(0, kHd), id_out = make_idesc(0, kC);
      const uint64_t dRA = make_smem_desc(smem_u32(RA)), dRB = make_smem_desc(smem_u32(RB)), dRC = make_smem_desc(smem_u32(RC)),
                     dRD = make_smem_desc(smem_u32(RD)), dRE = make_smem_desc(smem_u32(RE)), dRing = make_smem_desc(smem_u32(ring));
      // SWIZZLE_64B K-major descriptors of the mask operands: rows of 64 B, 8-row groups 512 B apart
      auto desc64 = [](uint32_t saddr) -> uint64_t {
        return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) |
               ((uint64_t)4 << 61);
      };
      const uint64_t dMA = desc64(smem_u32(MA)), dMB = desc64(smem_u32(MB));
      constexpr uint64_t CH = kChunk >> 4, VCH = kVtChunk >> 4, SL16 = kSlot >> 4;      // descriptor units (16 B)
      auto wslot = [&]() -> uint64_t {             // wait for the next weight block, return its descriptor
        const uint32_t slot = wcnt % kSlots, use = wcnt / kSlots;
#ifdef LFSR_DEBUG_HOOKS
        { const long long t0_ = clock64(); mbar_wait(bars + W_FULL + slot, use & 1u);
          if (p.dbg) { const long long dt_ = clock64() - t0_; p.dbg[blockIdx.x * 64 + 32] += dt_; p.dbg[blockIdx.x * 64 + 44 + (wcnt % kBlocks)] += dt_; } }
#else
        mbar_wait(bars + W_FULL + slot, use & 1u);
#endif
        tc_fence_after();
        return dRing + (uint64_t)slot * SL16;
      };
      auto wfree = [&]() { umma_commit(bars + W_EMPTY + (wcnt % kSlots)); ++wcnt; };
      const int pv_steps = p.nkeys >> 4;
#ifdef LFSR_DEBUG_HOOKS
      const long long mma_t0 = clock64();
#endif
      for (int it = 0; it < n_my; ++it) {
        const uint32_t par = (uint32_t)it & 1u;
        // ---- X = x W_in^T -> R (K = 64 input channels: one chunk)
        BT_MWAIT(1, bars + X_FULL, par);
        tc_fence_after();
        { const uint64_t b = wslot(); mma_chunk<false, true>(tR, dRA, b, id_f16); wfree(); }
        umma_commit(bars + R_FULL);
        // ---- V^T = Wv X^T -> T1 (as soon as the fp16 copy of X is written, while the epilogue still normalises)
        BT_MWAIT(2, bars + X_READY, par);
        tc_fence_after();
        { const uint64_t a = wslot(); mma_chunk<false, true>(tT1, a, dRA, id_f16); wfree(); }
        { const uint64_t a = wslot(); mma_chunk<false, false>(tT1, a, dRA + CH, id_f16); wfree(); }
        umma_commit(bars + VT_FULL);
        // ---- Q = N Wq^T, K = N Wk^T -> T2
        BT_MWAIT(3, bars + N_READY, par);
        tc_fence_after();
        { const uint64_t b = wslot(); mma_chunk<false, true>(tT2, dRB, b, id_f16); wfree(); }
        { const uint64_t b = wslot(); mma_chunk<false, false>(tT2, dRB + CH, b, id_f16); wfree(); }
        { const uint64_t b = wslot(); mma_chunk<false, true>(tT2 + 128, dRB, b, id_f16); wfree(); }
        { const uint64_t b = wslot(); mma_chunk<false, false>(tT2 + 128, dRB + CH, b, id_f16); wfree(); }
        umma_commit(bars + K_FULL);
        // ---- attention: S_h = Q_h K_h^T (one K = 16 MMA) into one of THREE score buffers, so the next head of each warp
        // pair is always ready; softmax by the epilogue; O_h = P_h V_h into the 16-column O buffer of its warp pair
        BT_MWAIT(4, bars + QKV_READY, par);
        tc_fence_after();
        auto issue_s = [&](int h) {
          const int b = h % 3;
          const uint64_t off = (uint64_t)(h >> 2) * CH + (uint64_t)(2 * (h & 3));
          const uint32_t ts = b == 2 ? tT1 : tT2 + 112 * b;
          umma_f16<0>(ts, dRC + off, dRD + off, id_s);
          umma_f16<1>(ts, dMA, dMB, id_s);
          umma_f16<1>(ts, dMA + 2, dMB + 2, id_s);
          umma_commit(bars + S_FULL + b);
        };
        issue_s(0);
        issue_s(1);
        issue_s(2);
        // X2 = X + O Wo^T is accumulated into R head by head while the attention runs: O_h (written into the slot of the
        // dead Q_h by the warps that later arrive on P_FULL of head h + 2) times the 16 matching K-columns of Wo
        uint64_t wo = 0;
        auto issue_wo = [&](int hh) {
          const uint64_t ko = (uint64_t)(2 * (hh & 3));
          umma_f16<1>(tR, dRC + (uint64_t)(hh >> 2) * CH + ko, wo + ko, id_f16);
        };
        for (int h = 0; h < kHeads; ++h) {
          const int j = h & 1;
          BT_MWAIT(5, bars + P_FULL + j, (uint32_t)(h >> 1) & 1u);
          tc_fence_after();
          const uint64_t pa = j ? dRA : dRB;
          const uint64_t vb = dRE + (uint64_t)(h * kHd * 128 >> 4);      // rows 16h.. of V^T (N = 16)
          for (int ks = 0; ks < pv_steps; ++ks) {
            const uint64_t aoff = (uint64_t)(ks >> 2) * CH + (uint64_t)(2 * (ks & 3));
            const uint64_t boff = (uint64_t)(ks >> 2) * VCH + (uint64_t)(2 * (ks & 3));
            if (ks == 0) umma_f16<0>(tT2 + 224 + 16 * j, pa + aoff, vb + boff, id_pv);
            else umma_f16<1>(tT2 + 224 + 16 * j, pa + aoff, vb + boff, id_pv);
          }
          umma_commit(bars + P_EMPTY + j);
          if (h + 3 < kHeads) issue_s(h + 3);           // its buffer was read by the softmax of head h (done: P_h is full)
          if (h >= 2) {
            if (h == 2 || h == 6) wo = wslot();          // Wo K-chunk 0 (heads 0..3) / 1 (heads 4..7)
            issue_wo(h - 2);
            if (h == 5) wfree();
          }
        }
        umma_commit(bars + X_EMPTY);
        BT_MWAIT(6, bars + O_READY, par);
        tc_fence_after();
        issue_wo(6);
        issue_wo(7);
        wfree();
        umma_commit(bars + R_FULL);
        // ---- F = N2 W1^T -> T2 (two halves of 128 columns)
        BT_MWAIT(7, bars + N2_READY, par);
        tc_fence_after();
        for (int n = 0; n < 2; ++n) {
          { const uint64_t b = wslot(); mma_chunk<false, true>(tT2 + 128 * n, dRB, b, id_f16); wfree(); }
          { const uint64_t b = wslot(); mma_chunk<false, false>(tT2 + 128 * n, dRB + CH, b, id_f16); wfree(); }
          umma_commit(bars + F_FULL + n);
        }
        // ---- X3 = X2 + relu(F) W2^T (accumulate into R); K = 256 in four chunks (RD, RD + chunk, RE, RE + chunk)
        for (int n = 0; n < 2; ++n) {
          BT_MWAIT(8, bars + F_READY + n, par);
          tc_fence_after();
          const uint64_t fa = n ? dRE : dRD;
          { const uint64_t b = wslot(); mma_chunk<false, false>(tR, fa, b, id_f16); wfree(); }
          { const uint64_t b = wslot(); mma_chunk<false, false>(tR, fa + CH, b, id_f16); wfree(); }
        }
        umma_commit(bars + R_FULL);
        // ---- y = X3 Wout^T -> T1[0..63]; the slot holds both K-chunks of Wout (64 rows x 128 B each)
        BT_MWAIT(9, bars + X3_READY, par);
        tc_fence_after();
        {
          const uint64_t b = wslot();
          mma_chunk<false, true>(tT1, dRB, b, id_out);
          mma_chunk<false, false>(tT1, dRB + CH, b + (uint64_t)(8192 >> 4), id_out);
          wfree();
        }
        umma_commit(bars + Y_FULL);
      }
#ifdef LFSR_DEBUG_HOOKS
      if (p.dbg) { p.dbg[blockIdx.x * 64 + 32 + 15] = clock64() - mma_t0; p.dbg[blockIdx.x * 64 + 32 + 14] = n_my; }
#endif
    }
    __syncwarp();
  } else {
    // ================= epilogue warps: TMEM -> registers -> fp16 operands in shared memory / y in HBM =================
    // four warps per TMEM lane quarter: `part` splits columns (LN, drains); for the softmax the quarter's warps form two
    // pairs (pair = head parity) whose two warps (sub) split the key columns of a head
    const int q = warp & 3, part = warp >> 2, pair = part >> 1, sub = part & 1;
    const int r = q * 32 + lane;                       // TMEM lane = tile row (token, s-major) - or channel for V^T
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const bool row_ok = r < p.nrows, row_st = r < kRows;      // row_st: the row exists in the 112-row operand buffers
    const int s_loc = r / p.A, a_idx = r - s_loc * p.A;
    // valid key columns of this query row: all A rows of the positions |s - s'| <= w that the tile holds
    const int k_lo = (s_loc - p.w > 0 ? s_loc - p.w : 0) * p.A;
    const int k_hi = ((s_loc + p.w < p.SL - 1 ? s_loc + p.w : p.SL - 1) + 1) * p.A;
    const int wlo = __reduce_min_sync(0xffffffffu, row_ok ? k_lo : 0x7fffffff);
    const int whi = __reduce_max_sync(0xffffffffu, row_ok ? k_hi : 0);
    // 16-column groups of the warp's key window, split between the two warps of the pair (<= 4 groups each)
    const int ngrp = p.nkeys >> 4;
    int g_first = whi > wlo ? wlo >> 4 : 0, g_end = whi > wlo ? (whi + 15) >> 4 : 0;
    const int g_mid = g_first + ((g_end - g_first + 1) >> 1);
    const int g0 = sub ? g_mid : g_first, g1 = sub ? g_end : g_mid;          // this warp: groups [g0, g1)
    const int z0 = sub ? g_end : 0, z1 = sub ? ngrp : g_first;               // groups it zero-fills in P
    const uint32_t sRA = smem_u32(RA), sRB = smem_u32(RB), sRC = smem_u32(RC), sRD = smem_u32(RD), sRE = smem_u32(RE);
    uint32_t rcnt = 0;
    auto arrive = [&](int bar) {          // generic-proxy writes -> async proxy (UMMA), TMEM reads ordered before the handoff
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + bar);
    };
    // LayerNorm over the 128 columns of R: this warp owns 32 of them, the other three warps of the lane quarter the rest
    // Each warp reduces its 32 columns to (mean_i, M2_i) locally; the four partials of a row are merged like Welford / Chan
    // (mean = sum mean_i / 4, M2 = sum M2_i + 32 sum (mean_i - mean)^2): numerically as robust as two passes over the row,
    // with one exchange through shared memory and one named barrier instead of four.
    auto layer_norm = [&](float* v, const float* g, const float* bta, float eps) {
      f2 acc = f2_pack(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 32; i += 2) acc = f2_add(acc, f2_pack(v[i], v[i + 1]));
      float s0, s1;
      f2_unpack(acc, s0, s1);
      const float mloc = (s0 + s1) * (1.f / 32.f);
      const f2 nml = f2_pack(-mloc, -mloc);
      acc = f2_pack(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const f2 d = f2_add(f2_pack(v[2 * i], v[2 * i + 1]), nml);
        acc = f2_fma(d, d, acc);
      }
      f2_unpack(acc, s0, s1);
      const int rx = row_st ? r : 0;                       // rows >= 112 are not tokens: they alias row 0's slot (never read back)
      if (row_st) reinterpret_cast<float2*>(xchg)[part * kRows + rx] = make_float2(mloc, s0 + s1);
      named_bar_sync(1 + q, 128);
      const float2 e0 = reinterpret_cast<const float2*>(xchg)[rx], e1 = reinterpret_cast<const float2*>(xchg)[kRows + rx];
      const float2 e2 = reinterpret_cast<const float2*>(xchg)[2 * kRows + rx], e3 = reinterpret_cast<const float2*>(xchg)[3 * kRows + rx];
      const float mean = (e0.x + e1.x + e2.x + e3.x) * 0.25f;
      const float d0 = e0.x - mean, d1 = e1.x - mean, d2 = e2.x - mean, d3 = e3.x - mean;
      const float m2 = (e0.y + e1.y + e2.y + e3.y) + 32.f * (d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3);
      const float rstd = rsqrtf(m2 * (1.f / kE) + eps);
      const f2 rs = f2_pack(rstd, rstd), nmr = f2_pack(-mean * rstd, -mean * rstd);
      const f2* g2 = reinterpret_cast<const f2*>(g + 32 * part);
      const f2* b2 = reinterpret_cast<const f2*>(bta + 32 * part);
#pragma unroll
      for (int i = 0; i < 16; ++i)
        f2_unpack(f2_fma(f2_fma(f2_pack(v[2 * i], v[2 * i + 1]), rs, nmr), g2[i], b2[i]), v[2 * i], v[2 * i + 1]);
    };
#ifdef LFSR_DEBUG_HOOKS
    long long t_prev = clock64();
    const bool prober = p.dbg && warp == 0 && lane == 0;
#define BT_TICK(i) do { if (prober) { const long long t_ = clock64(); p.dbg[blockIdx.x * 64 + (i)] += t_ - t_prev; t_prev = t_; } } while (0)
#else
#define BT_TICK(i) do { } while (0)
#endif
    const int cchunk = part >> 1, cu0 = 4 * (part & 1);   // this warp's 32 columns of a 128-column operand: chunk, first unit

    for (int it = 0; it < n_my; ++it) {
      const uint32_t par = (uint32_t)it & 1u;
      const int tile = (int)blockIdx.x + it * (int)gridDim.x;
      // ---- X -> fp16 X (B operand of the V^T GEMM), then N = LN1(X) (A operand of the Q / K GEMMs)
      mbar_wait(bars + R_FULL, rcnt++ & 1u);
      BT_TICK(0);
      tc_fence_after();
      {
        float v[32];
        tmem_ld32(tR + lane_off + 32 * part, v);
        tmem_wait_ld();
        if (row_st) { if (row_ok) store_f16<32>(sRA + cchunk * kChunk, r, cu0, v); else store_zero<32>(sRA + cchunk * kChunk, r, cu0); }
        arrive(X_READY);
        if (warp == 0 && lane == 0) bulk_wait_read0();      // the y store of the previous tile has finished reading RC
        layer_norm(v, p.ln[0], p.ln[1], p.eps1);
        if (row_st) { if (row_ok) store_f16<32>(sRB + cchunk * kChunk, r, cu0, v); else store_zero<32>(sRB + cchunk * kChunk, r, cu0); }
      }
      arrive(N_READY);
      BT_TICK(1);
      // ---- V^T rows (lane = channel, columns = tokens) -> RE ; Q (scaled, exp2 domain) -> RC ; K -> RD
      mbar_wait(bars + VT_FULL, par);
      BT_TICK(2);
      tc_fence_after();
      {
        float v[32];
        tmem_ld32(tT1 + lane_off + 32 * part, v);
        tmem_wait_ld();
        store_f16<32>(sRE + cchunk * kVtChunk, r, cu0, v);
      }
      BT_TICK(3);
      BT_TICK(4);
      BT_TICK(5);
      mbar_wait(bars + K_FULL, par);
      BT_TICK(6);
      tc_fence_after();
      {
        float v[32], u[32];
        tmem_ld32(tT2 + lane_off + 32 * part, v);                 // Q (1/sqrt(d) and log2 e are folded into Wq)
        tmem_ld32(tT2 + lane_off + 128 + 32 * part, u);           // K
        tmem_wait_ld();
        if (row_st) { store_f16<32>(sRC + cchunk * kChunk, r, cu0, v); store_f16<32>(sRD + cchunk * kChunk, r, cu0, u); }
      }
      arrive(QKV_READY);
      BT_TICK(7);
      // ---- softmax of this pair's heads (pair, pair + 2, ..): S in TMEM -> un-normalised exp2 as fp16 P in shared memory;
      // the O columns of the pair's previous head are drained (scaled by 1 / rowsum) as soon as its P.V has retired
      const uint32_t sP = pair ? sRA : sRB;
      // partial row maxima / sums of the pair's two warps: [pair][sub][128] each (the maxima reuse the LayerNorm exchange area)
      float* const pmax = xchg + pair * 2 * kRows;
      float* const psum = xchg + 4 * kRows + pair * 2 * kRows;
      const int rx = row_st ? r : 0;
      auto drain_o = [&](int h) {                         // 8 of the 16 columns of O_h per warp of the pair
        float v[8];
        {
          uint32_t t[8];
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                       : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7])
                       : "r"(tT2 + lane_off + 224 + 16 * pair + 8 * sub) : "memory");
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(t[i]);
        }
        const float inv = __fdividef(1.f, psum[rx] + psum[kRows + rx]);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] *= inv;
        if (row_st) store_f16<8>(sRC + (h >> 2) * kChunk, r, 2 * (h & 3) + sub, v);
      };
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        const int h = 2 * k + pair, b = h % 3;
        // completions of score buffer b before this head: (3, 3, 2) per tile for b = (0, 1, 2)
        BT_TICK(8);
        mbar_wait(bars + S_FULL + b, (uint32_t)(it * (b == 2 ? 2 : 3) + h / 3) & 1u);
        BT_TICK(9);
        tc_fence_after();
        const uint32_t tS = (b == 2 ? tT1 : tT2 + 112 * b) + lane_off;
        // pass 1: row maximum over this warp's key groups (the band mask is already in the scores: keys outside the band sit
        // ~30000 below the row maximum). The scores are NOT kept: pass 2 re-reads them from TMEM group by group, so that
        // only the packed fp16 exponentials (8 registers per group) stay live across the exchange and the P_EMPTY wait.
        float m0 = -1e30f, m1 = -1e30f;
        {
          float s[16 * kGrpPerWarp];
#pragma unroll
          for (int i = 0; i < kGrpPerWarp; ++i)       // (unconditional, clamped: keeps the array in registers)
            tmem_ld16(tS + (g0 + i < ngrp ? g0 + i : ngrp - 1) * 16, s + 16 * i);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < kGrpPerWarp; ++i)
            if (g0 + i < g1) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                m0 = fmaxf(m0, fmaxf(s[16 * i + j], s[16 * i + j + 1]));
                m1 = fmaxf(m1, fmaxf(s[16 * i + j + 2], s[16 * i + j + 3]));
              }
            }
        }
        if (row_st) pmax[sub * kRows + rx] = fmaxf(m0, m1);
        named_bar_sync(5 + 2 * q + pair, 64);              // (also: the partner has finished the previous head entirely)
        const float m = fmaxf(pmax[rx], pmax[kRows + rx]);
        const f2 negm = f2_pack(-m, -m);
        f2 acc = f2_pack(0.f, 0.f);
        uint32_t ph[8 * kGrpPerWarp];
#pragma unroll
        for (int i = 0; i < 8 * kGrpPerWarp; ++i) ph[i] = 0u;
#pragma unroll
        for (int i = 0; i < kGrpPerWarp; ++i)
          if (g0 + i < g1) {
            float s[16];
            tmem_ld16(tS + (g0 + i) * 16, s);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              float t0, t1;
              f2_unpack(f2_add(f2_pack(s[j], s[j + 1]), negm), t0, t1);
              const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
              ph[8 * i + (j >> 1)] = pack_f16x2(e0, e1);
              acc = f2_add(acc, f2_pack(e0, e1));
            }
          }
        float a0, a1;
        f2_unpack(acc, a0, a1);
        // only now is the P buffer needed: P.V of the pair's previous head has retired -> drain its O columns, reuse P / psum
        BT_TICK(10);
        mbar_wait(bars + P_EMPTY + pair, ((uint32_t)k & 1u) ^ 1u);
        BT_TICK(11);
        tc_fence_after();
        if (k > 0) drain_o(h - 2);
        named_bar_sync(5 + 2 * q + pair, 64);              // both warps have read psum / pmax of the previous round
#pragma unroll
        for (int i = 0; i < kGrpPerWarp; ++i)
          if (g0 + i < g1 && row_st) {
            const uint32_t rb = sP + (uint32_t)((g0 + i) >> 2) * kChunk + (uint32_t)r * 128u;
            const uint32_t u0 = (uint32_t)(2 * ((g0 + i) & 3)), sw = (uint32_t)(r & 7);
            st_shared_v4(rb + ((u0 ^ sw) << 4), ph[8 * i], ph[8 * i + 1], ph[8 * i + 2], ph[8 * i + 3]);
            st_shared_v4(rb + (((u0 + 1) ^ sw) << 4), ph[8 * i + 4], ph[8 * i + 5], ph[8 * i + 6], ph[8 * i + 7]);
          }
        if (k == 0)               // key groups outside this quarter's window are only ever written here: once per tile
          for (int g = z0; g < z1; ++g)
            if (row_st) store_zero<16>(sP + (g >> 2) * kChunk, r, 2 * (g & 3));
        if (row_st) psum[sub * kRows + rx] = a0 + a1;
        arrive(P_FULL + pair);
      }
      BT_TICK(8);
      mbar_wait(bars + P_EMPTY + pair, 1u);                // 5th wait of the tile: completion 4*it + 3
      BT_TICK(12);
      tc_fence_after();
      named_bar_sync(5 + 2 * q + pair, 64);                // the partner's last psum is visible
      drain_o(6 + pair);
      arrive(O_READY);
      BT_TICK(13);
      // ---- N2 = LN2(X2) -> RB
      mbar_wait(bars + R_FULL, rcnt++ & 1u);
      BT_TICK(14);
      tc_fence_after();
      {
        float v[32];
        tmem_ld32(tR + lane_off + 32 * part, v);
        tmem_wait_ld();
        layer_norm(v, p.ln[2], p.ln[3], p.eps2);
        if (row_st) store_f16<32>(sRB + cchunk * kChunk, r, cu0, v);
      }
      arrive(N2_READY);
      BT_TICK(15);
      // ---- relu(F): this warp's 64 of the 256 columns -> K-chunk `part` of the FFN operand (RD: 0, 1; RE: 2, 3)
      mbar_wait(bars + F_FULL + (part >> 1), par);
      BT_TICK(16);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld32(tT2 + lane_off + 64 * part + 32 * c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        if (row_st) store_f16<32>((part >> 1 ? sRE : sRD) + (part & 1) * kChunk, r, 4 * c, v);
      }
      arrive(F_READY + (part >> 1));
      BT_TICK(17);
      // ---- X3 -> fp16 in RB (A operand of linear_out)
      mbar_wait(bars + R_FULL, rcnt++ & 1u);
      BT_TICK(18);
      tc_fence_after();
      {
        float v[32];
        tmem_ld32(tR + lane_off + 32 * part, v);
        tmem_wait_ld();
        if (row_st) store_f16<32>(sRB + cchunk * kChunk, r, cu0, v);
      }
      arrive(X3_READY);
      BT_TICK(19);
      // ---- y rows of the tile's own query positions: staged as fp32 [rows][32] x 2 chunks (SWIZZLE_128B) in RC (O is dead),
      // then ONE lane hands the two boxes to the TMA unit (same 5-D addressing as the input tile, clipped to the tensor)
      mbar_wait(bars + Y_FULL, par);
      BT_TICK(20);
      tc_fence_after();
      {
        float v[16];
        tmem_ld16(tT1 + lane_off + 16 * part, v);
        tmem_wait_ld();
        const int seq = tile / p.nt, t = tile - seq * p.nt;
        const int b = seq / p.npq, pq = seq - b * p.npq;
        int ls0 = t * p.sq - p.w;
        if (ls0 < 0) ls0 = 0;
        if (ls0 > p.S - p.SL) ls0 = p.S - p.SL;
        const int qs0 = t * p.sq;
        const int rq = r - (qs0 - ls0) * p.A;                 // row inside the query block
        if (rq >= 0 && rq < p.sq * p.A) store_f16<16>(sRC, rq, 2 * part, v);
        fence_proxy_async();
        tc_fence_before();
        named_bar_sync(13, 32 * kEpiWarps);
        if (warp == 0 && lane == 0) {
          tma_store_5d(&tmY, RC, 0, 0, qs0, pq, b);
          bulk_commit();
        }
      }
      tc_fence_before();
      BT_TICK(21);
    }
  }
  if (warp == 0 && lane == 0) bulk_wait0();                 // all y stores have landed
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static float round_tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return x;
  u += 0xFFFu + ((u >> 13) & 1u);
  u &= ~0x1FFFu;
  float y;
  memcpy(&y, &u, 4);
  return y;
}

// tile plan: nt tiles per sequence, each owns sq query positions and loads SL positions (SL * A <= 112 rows)
static bool plan(int A, int S, int w, int* nt_, int* sq_, int* SL_) {
  for (int nt = 1; nt <= S; ++nt) {
    const int sq = (S + nt - 1) / nt;
    int SL = 0;
    for (int t = 0; t * sq < S; ++t) {
      const int q0 = t * sq, q1 = q0 + sq < S ? q0 + sq : S;
      const int l0 = q0 - w > 0 ? q0 - w : 0, l1 = q1 + w < S ? q1 + w : S;
      if (l1 - l0 > SL) SL = l1 - l0;
    }
    if (SL * A <= kMaxKeys) { *nt_ = (S + sq - 1) / sq; *sq_ = sq; *SL_ = SL; return true; }
  }
  return false;
}

static bool geometry_ok(const lfsr_tensor* x, const lfsr_tensor* y, const lfsr_basictrans_desc* d) {
  if (!tensor_ok(x) || !tensor_ok(y) || !d) return false;
  if (d->E != kE || d->C != kC || d->heads != kHeads) return false;
  if (x->c != kC || y->c != kC || x->ld % 8 || y->ld % 8 || ((uintptr_t)x->ptr & 15) || ((uintptr_t)y->ptr & 15)) return false;
  if (x->n != y->n || x->h != y->h || x->w != y->w) return false;
  if (d->A < 1 || d->S < 1 || d->half_window < 0 || d->nb < 1 || d->np < 1 || d->nq < 1) return false;
  if (d->stride_p != (int64_t)d->nq * d->stride_q) return false;          // (p, q) must merge into one tensor-map dimension
  if (d->A > 256 || d->S > 4096) return false;
  int nt, sq, SL;
  if (!plan(d->A, d->S, d->half_window, &nt, &sq, &SL)) return false;
  if ((long long)d->nb * d->np * d->nq * nt > 0x7fffffffLL) return false;
  // the two softmax warps of a pair keep <= kGrpPerWarp 16-column groups of their quarter's key window in registers
  for (int q = 0; q < 4; ++q) {
    int lo = 1 << 30, hi = 0;
    for (int r = 32 * q; r < 32 * q + 32 && r < SL * d->A; ++r) {
      const int sl = r / d->A, w = d->half_window;
      const int kl = (sl - w > 0 ? sl - w : 0) * d->A, kh = ((sl + w < SL - 1 ? sl + w : SL - 1) + 1) * d->A;
      lo = kl < lo ? kl : lo; hi = kh > hi ? kh : hi;
    }
    if (hi > lo && ((hi + 15) / 16 - lo / 16 + 1) / 2 > kGrpPerWarp) return false;
  }
  return true;
}

}  // namespace bt
}  // namespace lfsr

using namespace lfsr;
using namespace lfsr::bt;

extern "C" size_t lfsr_basictrans_packed_bytes(void) { return (size_t)kBlocks * kSlot; }

// Weights in torch layout [out][in] (fp32, host): linear_in [128][64], MultiheadAttention.in_proj_weight [384][128]
// (q | k | v), out_proj [128][128], feed_forward.1 [256][128], feed_forward.4 [128][256], linear_out [64][128] ->
// 19 blocks of [128 rows][128 bytes] in the order the kernel consumes them (see kW* above). linear_in stays fp32 rounded
// to TF32 (its A operand is the TMA-loaded input tile); everything else is fp16.
extern "C" int lfsr_pack_basictrans(const float* w_in, const float* w_qkv, const float* w_o, const float* w_ff1,
                                    const float* w_ff2, const float* w_out, void* packed_host) {
  LFSR_REQUIRE(w_in && w_qkv && w_o && w_ff1 && w_ff2 && w_out && packed_host, "lfsr_pack_basictrans: null pointer");
  uint8_t* out = static_cast<uint8_t*>(packed_host);
  memset(out, 0, (size_t)kBlocks * kSlot);
  auto blk_f16 = [&](int blk) { return reinterpret_cast<__half*>(out + (size_t)blk * kSlot); };
  auto pack_f16 = [&](int blk0, const float* w, int row0, int rows, int in_dim, int chunks) {
    for (int c = 0; c < chunks; ++c)
      for (int o = 0; o < rows; ++o)
        for (int k = 0; k < 64; ++k)
          blk_f16(blk0 + c)[o * 64 + k] = __float2half_rn(w[(size_t)(row0 + o) * in_dim + 64 * c + k]);
  };
  pack_f16(kWin, w_in, 0, kE, kC, 1);
  pack_f16(kWv, w_qkv, 2 * kE, kE, kE, 2);
  {   // Wq with 1/sqrt(head_dim) (MultiheadAttention scales q) and log2(e) (the softmax runs on exp2) folded in
    const float qscale = 1.4426950408889634f / sqrtf((float)kHd);
    for (int c = 0; c < 2; ++c)
      for (int o = 0; o < kE; ++o)
        for (int k = 0; k < 64; ++k)
          blk_f16(kWq + c)[o * 64 + k] = __float2half_rn(w_qkv[(size_t)o * kE + 64 * c + k] * qscale);
  }
  pack_f16(kWk, w_qkv, kE, kE, kE, 2);
  pack_f16(kWo, w_o, 0, kE, kE, 2);
  pack_f16(kW1, w_ff1, 0, kE, kE, 2);
  pack_f16(kW1 + 2, w_ff1, kE, kE, kE, 2);
  pack_f16(kW2, w_ff2, 0, kE, 2 * kE, 4);
  for (int c = 0; c < 2; ++c)                           // W_out: both K-chunks in one slot, 64 rows x 128 B each
    for (int o = 0; o < kC; ++o)
      for (int k = 0; k < 64; ++k)
        blk_f16(kWout)[(size_t)c * 4096 + o * 64 + k] = __float2half_rn(w_out[(size_t)o * kE + 64 * c + k]);
  return LFSR_OK;
}

extern "C" int lfsr_basictrans_supported(const lfsr_tensor* x, const lfsr_tensor* y, const lfsr_basictrans_desc* d) {
  return geometry_ok(x, y, d) && get_encode() != nullptr ? 1 : 0;
}

extern "C" int lfsr_epit_basictrans(const lfsr_tensor* x, const void* packed, const lfsr_tensor* y,
                                    const lfsr_basictrans_desc* d, void* stream) {
  LFSR_REQUIRE(packed, "lfsr_epit_basictrans: null weights");
  LFSR_REQUIRE(geometry_ok(x, y, d), "lfsr_epit_basictrans: unsupported geometry (query lfsr_basictrans_supported)");
  LFSR_REQUIRE(x->ptr != y->ptr, "lfsr_epit_basictrans: in-place operation is not supported");
  EncodeTiledFn encode = get_encode();
  if (!encode) { set_error("lfsr_epit_basictrans: cuTensorMapEncodeTiled unavailable"); return LFSR_ERR_CUDA; }
  Params p;
  memset(&p, 0, sizeof(p));
  p.A = d->A; p.S = d->S; p.w = d->half_window;
  plan(p.A, p.S, p.w, &p.nt, &p.sq, &p.SL);
  p.nrows = p.SL * p.A;
  p.nkeys = (p.nrows + 15) / 16 * 16;
  p.npq = d->np * d->nq; p.nq = d->nq;
  p.nseq = d->nb * p.npq;
  p.total_tiles = p.nseq * p.nt;
  p.stride_a = d->stride_a; p.stride_s = d->stride_s; p.stride_b = d->stride_b; p.stride_p = d->stride_p; p.stride_q = d->stride_q;
  p.y = (float*)y->ptr; p.ld_y = y->ld;
  p.eps1 = d->eps1; p.eps2 = d->eps2;
  p.qscale = 1.4426950408889634f / sqrtf((float)kHd);
  p.dbg = dbg_env("LFSR_BT_DBG_PTR") ? (long long*)strtoull(dbg_env("LFSR_BT_DBG_PTR"), nullptr, 0) : nullptr;
  memcpy(p.ln[0], d->ln1_g, sizeof(float) * kE); memcpy(p.ln[1], d->ln1_b, sizeof(float) * kE);
  memcpy(p.ln[2], d->ln2_g, sizeof(float) * kE); memcpy(p.ln[3], d->ln2_b, sizeof(float) * kE);
  const long long T = (long long)x->n * x->h * x->w;
  // every token the tiles touch must lie inside the tensor
  const long long last = (long long)(d->nb - 1) * d->stride_b + (long long)(d->np - 1) * d->stride_p +
                         (long long)(d->nq - 1) * d->stride_q + (long long)(d->A - 1) * d->stride_a +
                         (long long)(d->S - 1) * d->stride_s;
  LFSR_REQUIRE(last < T, "lfsr_epit_basictrans: the stride set addresses tokens outside the tensor");
  CUtensorMap tmX, tmW;
  {
    const cuuint64_t ld_b = (cuuint64_t)x->ld * 2;
    cuuint64_t dims[5] = {(cuuint64_t)kC, (cuuint64_t)p.A, (cuuint64_t)p.S, (cuuint64_t)p.npq, (cuuint64_t)d->nb};
    cuuint64_t strides[4] = {ld_b * (cuuint64_t)d->stride_a, ld_b * (cuuint64_t)d->stride_s, ld_b * (cuuint64_t)d->stride_q,
                             ld_b * (cuuint64_t)d->stride_b};
    cuuint32_t box[5] = {64, (cuuint32_t)p.A, (cuuint32_t)p.SL, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, x->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("lfsr_epit_basictrans: cuTensorMapEncodeTiled(x) failed with %d", (int)r); return LFSR_ERR_CUDA; }
  }
  {
    cuuint64_t dims[2] = {128, (cuuint64_t)kBlocks * 128};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {128, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(packed), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("lfsr_epit_basictrans: cuTensorMapEncodeTiled(w) failed with %d", (int)r); return LFSR_ERR_CUDA; }
  }
  CUtensorMap tmY;
  {
    const cuuint64_t ld_b = (cuuint64_t)y->ld * 2;
    cuuint64_t dims[5] = {(cuuint64_t)kC, (cuuint64_t)p.A, (cuuint64_t)p.S, (cuuint64_t)p.npq, (cuuint64_t)d->nb};
    cuuint64_t strides[4] = {ld_b * (cuuint64_t)d->stride_a, ld_b * (cuuint64_t)d->stride_s, ld_b * (cuuint64_t)d->stride_q,
                             ld_b * (cuuint64_t)d->stride_b};
    cuuint32_t box[5] = {64, (cuuint32_t)p.A, (cuuint32_t)p.sq, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&tmY, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, y->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("lfsr_epit_basictrans: cuTensorMapEncodeTiled(y) failed with %d", (int)r); return LFSR_ERR_CUDA; }
  }
  const int smem = kSmemBytes + 1024;
  static DevOnce once;
  if (once.need()) {
    if (opt_in_smem(basictrans_kernel, smem, "lfsr_epit_basictrans")) return LFSR_ERR_CUDA;
    once.done();
  }
  const int sms = sm_count_current();
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  basictrans_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmX, tmW, tmY, p);
  return check_launch("basictrans_kernel");
}
