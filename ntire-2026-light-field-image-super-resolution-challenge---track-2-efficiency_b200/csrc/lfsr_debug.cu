// Hardware probe (not on the product path): does a K-major SWIZZLE_128B UMMA operand descriptor
// accept a start address that is a multiple of 128 B but not of 1024 B (row-shifted view into a
// larger TMA-written tile)? Needed for halo-resident 3x3 convolutions. D[m][n] = A[m + shift][n]
// with B = identity; variant 0 leaves base_offset = 0, variant 1 sets base_offset = (addr >> 7) & 7.
#include <cuda.h>
#include <string.h>
#include "lfsr_common.cuh"

namespace lfsr { namespace dbg {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int shift, int variant) {
  extern __shared__ uint8_t raw[];
  const uint32_t r0 = smem_u32(raw);
  uint8_t* smem = raw + (((r0 + 1023u) & ~1023u) - r0);
  uint8_t* sA = smem;                 // 256 rows x 128 B
  uint8_t* sB = smem + 256 * 128;     // 32 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32 * 128);
  uint64_t* bar2 = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar2 + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(256 * 128 + 32 * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sA)), "l"(&tmA), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sB)), "l"(&tmB), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
  }
  // everyone waits for the data
  {
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 24) && !done; ++it)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)) : "memory");
    if (!done) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_addr = smem_u32(sA) + shift * 128 + k * 32;
      const uint32_t b_addr = smem_u32(sB) + k * 32;
      uint64_t da = ((uint64_t)((a_addr & 0x3FFFFu) >> 4)) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      if (variant == 1) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
      const uint64_t db = ((uint64_t)((b_addr & 0x3FFFFu) >> 4)) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(k != 0)) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar2)) : "memory");
  }
  {
    uint32_t done = 0;
    for (uint32_t it = 0; it < (1u << 24) && !done; ++it)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar2)) : "memory");
    if (!done) __trap();
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[32];
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                 "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr + 16) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}
}}  // namespace

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// a_dev: [256][32] floats, b_dev: [32][32] floats, out_dev: [128][32] floats
extern "C" int lfsr_debug_umma_shift(const float* a_dev, const float* b_dev, float* out_dev, int shift, int variant, void* stream) {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) return -2;
  EncodeTiledFn encode = (EncodeTiledFn)fp;
  CUtensorMap tmA, tmB;
  cuuint64_t dimsA[2] = {32, 256}, dimsB[2] = {32, 32}, strides[1] = {128};
  cuuint32_t boxA[2] = {32, 256}, boxB[2] = {32, 32}, es[2] = {1, 1};
  if (encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)a_dev, dimsA, strides, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -3;
  if (encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)b_dev, dimsB, strides, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return -3;
  const size_t smem = 1024 + 256 * 128 + 32 * 128 + 64;
  cudaFuncSetAttribute(lfsr::dbg::probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lfsr::dbg::probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmB, out_dev, shift, variant);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- probe 2: issue-to-completion time of a chain of tf32 UMMAs (cycles), same vs alternating accumulators,
// K-advance inside one 128 B swizzle atom (kstep 32 B) vs one MMA per fresh atom row block.
namespace lfsr { namespace dbg {
__global__ void __launch_bounds__(128, 1)
mma_rate_kernel(long long* out, int N, int chain, int alternate, int kmode) {
  extern __shared__ uint8_t raw[];
  const uint32_t r0 = smem_u32(raw);
  uint8_t* smem = raw + (((r0 + 1023u) & ~1023u) - r0);
  uint8_t* sA = smem;                  // 4 x 16 KB
  uint8_t* sB = smem + 4 * 16384;      // 4 x 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 4 * 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(bar + 2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const long long t0 = clock64();
    for (int i = 0; i < chain; ++i) {
      const int st = (i >> 2) & 3, k = i & 3;
      const uint32_t a_addr = smem_u32(sA) + st * 8192 + kmode * 128 + k * 32;   // kmode = row shift of the A view
      const uint32_t b_addr = smem_u32(sB) + st * 32768 + k * 32;
      const uint64_t da = ((uint64_t)((a_addr & 0x3FFFFu) >> 4)) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint64_t db = ((uint64_t)((b_addr & 0x3FFFFu) >> 4)) | ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t d = tmem + ((alternate == 1 && (i & 1)) ? 256 : 0);
      if (alternate == 16)       // same descriptors read as fp16 (K = 16 per instruction, same 32 B per row): cost per operand BYTE
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(da), "l"(db), "r"(idesc & ~((7u << 7) | (7u << 10))), "r"((uint32_t)(i > 1)) : "memory");
      else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)(i > 1)) : "memory");
      // alternate == 4 / 8: commit to a scratch mbarrier after every 4th / 8th MMA (cost of tcgen05.commit in the issue
      // stream; mask test instead of a modulo so the probe itself stays cheap)
      if ((alternate == 4 && (i & 3) == 3) || (alternate == 8 && (i & 7) == 7))
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 2)) : "memory");
    }
    const long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)) : "memory");
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
}}  // namespace

extern "C" int lfsr_debug_mma_rate(long long* out_dev, int n, int chain, int alternate, int kmode, void* stream) {
  const size_t smem = 1024 + 4 * 16384 + 4 * 32768 + 64;
  cudaFuncSetAttribute(lfsr::dbg::mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  lfsr::dbg::mma_rate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(out_dev, n, chain, alternate, kmode);
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

// ---- probe 3: the same chain issued the way conv_tc_kernel issues now - converged warp, elect.sync, descriptors advanced
// by uniform adds, four K-steps unrolled - so the ISSUE stream is not what is measured. KIND 0: kind::tf32 (K = 8),
// 1: kind::f16 over the same bytes (K = 16, same 32 B per row). afix / bfix: reuse the same A / B slice for every MMA.
namespace lfsr { namespace dbg {
template <int KIND>
__device__ __forceinline__ void mma_one(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
template <int KIND>
__global__ void __launch_bounds__(128, 1)
mma_rate2_kernel(long long* out, int N, int chain, int afix, int bfix) {
  extern __shared__ uint8_t raw[];
  const uint32_t r0 = smem_u32(raw);
  uint8_t* smem = raw + (((r0 + 1023u) & ~1023u) - r0);
  uint8_t* sA = smem;                  // 4 x 16 KB
  uint8_t* sB = smem + 4 * 16384;      // 4 x 32 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 4 * 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 4);
  for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 1.0f;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + 1)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1000000;" ::"r"(smem_u32(bar + 2)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar + 1)) : "memory");   // phase 0 of bar+1 is complete
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  if (warp == 0) {
    const uint32_t fmt = KIND == 0 ? 2u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t hi = ((uint64_t)1 << 16) | ((uint64_t)64 << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
    const uint64_t da0 = ((uint64_t)((smem_u32(sA) & 0x3FFFFu) >> 4)) | hi, db0 = ((uint64_t)((smem_u32(sB) & 0x3FFFFu) >> 4)) | hi;
    const uint32_t a_st = afix ? 0u : (16384u >> 4), b_st = bfix ? 0u : (32768u >> 4);
    const uint32_t a_k = afix ? 0u : 2u, b_k = bfix ? 0u : 2u;
    // bmode (afix bits 4..7): what separates consecutive groups of `grp` (afix bits 8..15, default 4) MMAs, as in the conv
    // kernel's stage loop: 1 = leave / re-enter the elect block (+ __syncwarp), 2 = + tcgen05.commit to a scratch barrier,
    // 3 = + a wait on an already completed mbarrier phase + tcgen05.fence::after_thread_sync
    const int bmode = (afix >> 4) & 15;
    const int grp = ((afix >> 8) & 255) ? ((afix >> 8) & 255) : 4;
    afix &= 1;
    const long long t0 = clock64();
    uint32_t pred;
    if (bmode) {
      for (int i = 0; i < chain; i += grp) {
        if (bmode >= 3) {
          uint32_t ok = 0;
          for (int spin = 0; !ok && spin < (1 << 20); ++spin)      // bounded: a probe must not hang the box
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(bar + 1)) : "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) {
          for (int j = i; j < i + grp; j += 4) {
            const uint32_t st = (uint32_t)(j >> 2) & 3u;
            const uint64_t da = da0 + st * a_st, db = db0 + st * b_st;
            mma_one<KIND>(tmem, da, db, idesc, (uint32_t)(j > 0));
            mma_one<KIND>(tmem, da + a_k, db + b_k, idesc, 1u);
            mma_one<KIND>(tmem, da + 2 * a_k, db + 2 * b_k, idesc, 1u);
            mma_one<KIND>(tmem, da + 3 * a_k, db + 3 * b_k, idesc, 1u);
          }
          if (bmode >= 2)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar + 2)) : "memory");
          if (i + grp >= chain)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        __syncwarp();
      }
    } else {
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    if (pred) {
      for (int i = 0; i < chain; i += 4) {
        const uint32_t st = (uint32_t)(i >> 2) & 3u;
        const uint64_t da = da0 + st * a_st, db = db0 + st * b_st;
        mma_one<KIND>(tmem, da, db, idesc, (uint32_t)(i > 0));
        mma_one<KIND>(tmem, da + a_k, db + b_k, idesc, 1u);
        mma_one<KIND>(tmem, da + 2 * a_k, db + 2 * b_k, idesc, 1u);
        mma_one<KIND>(tmem, da + 3 * a_k, db + 3 * b_k, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    }
    __syncwarp();
    const long long t1 = clock64();
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(bar)) : "memory");
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
}}  // namespace

extern "C" int lfsr_debug_mma_rate2(long long* out_dev, int n, int chain, int kind, int afix, int bfix, void* stream) {
  const size_t smem = 1024 + 4 * 16384 + 4 * 32768 + 64;
  if (kind == 0) {
    cudaFuncSetAttribute(lfsr::dbg::mma_rate2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lfsr::dbg::mma_rate2_kernel<0><<<1, 128, smem, (cudaStream_t)stream>>>(out_dev, n, chain, afix, bfix);
  } else {
    cudaFuncSetAttribute(lfsr::dbg::mma_rate2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    lfsr::dbg::mma_rate2_kernel<1><<<1, 128, smem, (cudaStream_t)stream>>>(out_dev, n, chain, afix, bfix);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
