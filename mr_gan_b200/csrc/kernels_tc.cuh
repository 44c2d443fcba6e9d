// kernels_tc.cuh -- tcgen05 / TMEM / TMA grouped GEMM of the training step (precision = tf32).
//
// One warp-specialised kernel serves all dense contractions of the step, grouped over folds
// (blockIdx.z) with per-(op, fold) TMA descriptors resident in HBM:
//
//   role            MMA-A (M = 128 "feature" rows)          MMA-B (N = bn columns)            accumulator
//   forward         W^T  from Waug[k, n]  (MN-major)        act  from A[r, k]   (K-major)     D[n, r]
//   backward dX     W    from Waug[k, n]  (K-major)         dZ   from dZ[r, n]  (K-major)     D[k, r]
//   backward dW     dZ^T from dZ[r, n]    (MN-major)        A^T  from A[r, k]   (MN-major)    D[n, k]
//
// i.e. the weight-side dimension always sits on the 128 TMEM lanes ("swap-AB": the batch is only
// 50..150 rows, mr_gan.py:78) and fp32 master weights feed kind::tf32 MMAs directly -- no low
// precision shadow copy, no transposed copy (both operand majors are legal for tf32 descriptors).
// The lane <-> memory-contiguous-dimension match makes every epilogue access coalesced:
// thread = one feature, registers = rows (forward / dX) or input features (dW).
//
// Pipeline: warp 0 = TMA producer (4-stage mbarrier ring, 32 contraction elements per stage),
// warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue
// (tcgen05.ld 32x32b -> bias-free activation + Philox GaussianNoise | act' mask | fused Adam).
#pragma once
#include <cuda.h>
#include "common.cuh"

enum { EPI_ADAM = 3 };

struct alignas(64) TcOp {
  CUtensorMap mapA, mapB;
  GemmDesc g;                 // epilogue description shared with the fp32 path (C, C2, aux, act, noise ...)
  float *P, *Mo, *Vo;         // fused Adam (EPI_ADAM): tensor base inside the flat buffers
  int ME, NE, KE;             // extents: MMA-M (features), MMA-N, contraction
  int bn;                     // MMA-N per tile (multiple of 16, <= 256; multiple of 32 when B is MN-major)
  int epi;                    // EPI_FWD / EPI_DX / EPI_STORE / EPI_ADAM
  int net;                    // 0 = discriminator, 1 = generator (which lr_t the fused Adam uses)
  int ws_stride;              // split-K (large-batch dW): floats between the partial products of two contraction slices
  float* ws;                  // ... and their workspace, laid out like C; k_splitk_reduce sums the slices in a fixed order
  int esz;                    // operand element size: 4 (or 0) = fp32 read as tf32, 2 = fp16 (kind::f16, fp32 accumulation)
  int afmt, bfmt;             // esz == 2: element format of the A / B operand, 0 = f16, 1 = bf16 (gradient-side operands, optional)
  int pad_[1];
};

#define TC_KBLK 32            // contraction elements per stage (one 128-byte swizzle row of fp32)
#define TC_SPIN_LIMIT (1ll << 31)   // cycles; a stuck barrier traps instead of hanging the GPU

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > TC_SPIN_LIMIT) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor (version 1).  lbo/sbo in bytes.
//   K-major operand : layout 2 = SWIZZLE_128B (16-byte chunks XOR row&7; TMA SWIZZLE_128B), sbo = 8 rows = 1024 B
//   MN-major operand: 32-bit types only support layout 1 = SWIZZLE_128B_BASE32B (32-byte chunks XOR k-row&3;
//                     TMA SWIZZLE_128B_ATOM_32B): 4 k-rows of 128 B per atom -> sbo = 512 B, lbo = next 32-element MN block
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float v[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float v[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float v[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float v[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace tc

// A_MN / B_MN: operand is MN-major (its MMA M/N dimension is the memory-contiguous one).
// STAGES: depth of the TMA->MMA ring; TMEM_COLS: accumulator columns allocated (power of 2 >= bn);
// MINB: CTAs per SM the register allocation must allow (dW uses 2 so that one CTA's Adam epilogue
// streams HBM while the other loads operands and runs its MMAs).
// EPW: epilogue warps (4 or 8; with 8 the accumulator columns are split between two warp groups).
// MT: feature sub-tiles of 128 per CTA (1 or 2).  With MT = 2 the CTA computes 256 features against ONE copy of the
// batch-side operand tile (two accumulators in TMEM, 2 x 4 MMAs per stage), which halves the activation bytes every
// SM has to ingest -- the forward / dX mainloops are bound by L2->SM traffic, not by HBM or the tensor pipe -- and each
// warp group of the epilogue owns one sub-tile.  MT = 2 requires EPW = 8 and TMEM_COLS = 512.
template <bool A_MN, bool B_MN, int STAGES, int TMEM_COLS, int MINB, int EPW, int MT>
__global__ void __launch_bounds__(64 + 32 * EPW, MINB)
k_gemm_tc(const TcOp* __restrict__ ops, FoldState* __restrict__ folds, int rows_override, AdamHyper hp, OperandMode om) {
  using namespace tc;
  pdl_launch_dependents();
  // dW (B_MN) kernels take the number of contraction slices in `rows_override`: blockIdx.z = fold * ksplit + slice, each
  // slice stores its partial product to op.ws (deterministic split-K: a fold's dW of the narrow layers is one or two
  // tiles, which would leave a large-batch contraction of thousands of rows on one or two SMs)
  const int ksplit = (B_MN && rows_override > 1) ? rows_override : 1;
  const int kslice = (int)blockIdx.z % ksplit;
  const TcOp& op = ops[blockIdx.z / ksplit];      // descriptor tables are written once at handle creation: safe before pdl_wait()
  const int ME = op.ME, KE = op.KE, bn = op.bn;
  int NE = op.NE;
  if (rows_override > 0 && !B_MN) NE = rows_override;      // fewer batch rows (G step)
  const int m0 = blockIdx.x * 128 * MT, n0 = blockIdx.y * bn;
  if (m0 >= ME || n0 >= NE) return;                         // uniform per CTA: nothing allocated yet

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t a_bytes = MT * 128 * 128, b_bytes = (uint32_t)bn * 128, stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * stage_bytes);
  uint64_t* full = bars;                    // [STAGES]
  uint64_t* empty = bars + STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 16-bit operands: a 128-byte swizzle row holds 64 contraction elements, an MMA covers K = 16, an MN-major box is
  // 64 elements x 64 contraction rows (plain 128B swizzle, UMMA layout 2, SBO = 1024 B, LBO = 8192 B)
  const bool f16 = op.esz == 2;
  const int kblk = f16 ? 64 : TC_KBLK;
  const int nkb_all = (KE + kblk - 1) / kblk;
  const int kb_per = (nkb_all + ksplit - 1) / ksplit;
  const int kb0 = kslice * kb_per;                           // the host picks ksplit so that no slice is empty
  const int nkb = min(nkb_all, kb0 + kb_per) - kb0;

  if (warp == 0 && lane == 0) {
    prefetch_map(&op.mapA);
    prefetch_map(&op.mapB);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: TMEM_COLS fp32 columns x 128 lanes for the accumulator tile
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"((uint32_t)TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();                             // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], stage_bytes);
        uint8_t* sa = smem + (size_t)s * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        const int k0 = (kb0 + kb) * kblk;
        const int bbytes = f16 ? 8192 : 4096;        // one MN-major box: kblk elements wide x kblk contraction rows
        if (A_MN) {
          for (int b = 0; b < (128 / kblk) * MT; ++b) tma_load_2d(&op.mapA, &full[s], sa + b * bbytes, m0 + kblk * b, k0);
        } else {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) tma_load_2d(&op.mapA, &full[s], sa + mt * 16384, k0, m0 + 128 * mt);
        }
        if (B_MN) {
          for (int b = 0; b < bn / kblk; ++b) tma_load_2d(&op.mapB, &full[s], sb + b * bbytes, n0 + kblk * b, k0);
        } else {
          tma_load_2d(&op.mapB, &full[s], sb, k0, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      // instruction descriptor operand formats: 0 = f16, 1 = bf16, 2 = tf32; D format 1 = f32
      const uint32_t fmtA = f16 ? (uint32_t)op.afmt : 2u, fmtB = f16 ? (uint32_t)op.bfmt : 2u;
      const uint32_t idesc = (1u << 4) | (fmtA << 7) | (fmtB << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(bn >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t mn_lbo = f16 ? 8192u : 4096u, mn_sbo = f16 ? 1024u : 512u, mn_lay = f16 ? 2u : 1u, mn_step = f16 ? 2048u : 1024u;
      const uint32_t lboA = A_MN ? mn_lbo : 16u, lboB = B_MN ? mn_lbo : 16u;
      const uint32_t sboA = A_MN ? mn_sbo : 1024u, sboB = B_MN ? mn_sbo : 1024u;
      const uint32_t layA = A_MN ? mn_lay : 2u, layB = B_MN ? mn_lay : 2u;
      const uint32_t stepA = A_MN ? mn_step : 32u, stepB = B_MN ? mn_step : 32u;   // bytes per MMA (8 tf32 / 16 f16 contraction elements)
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + a_bytes;
#pragma unroll
        for (int k = 0; k < TC_KBLK / 8; ++k) {
          const uint64_t db = smem_desc(sb + k * stepB, lboB, sboB, layB);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {      // sub-tile mt: A block at +16 KB, accumulator at TMEM column 256 * mt
            const uint64_t da = smem_desc(sa + mt * 16384u + k * stepA, lboA, sboA, layA);
            if (f16) mma_f16(tmem_base + 256u * mt, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            else mma_tf32(tmem_base + 256u * mt, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
        }
        mma_commit(&empty[s]);            // frees the smem stage when these MMAs retire
      }
      mma_commit(tmem_full);              // accumulator complete
    }
  } else {
    // ===================== epilogue (EPW warps; a warp may only touch TMEM lanes 32*(warp%4)..+31) =====================
    const int lane_base = 32 * (warp & 3);
    const int wg = (warp - 2) >> 2;                       // epilogue warp group (0 or 1 when EPW == 8)
    const int sub = (MT == 2) ? wg : 0;                   // MT == 2: one feature sub-tile per warp group
    const int chalf = (EPW == 8 && MT == 1) ? (((min(bn, NE - n0) + 1) / 2 + 15) & ~15) : bn;   // columns per warp group
    const int cbeg = (EPW == 8 && MT == 1) ? wg * chalf : 0;
    const int f = m0 + 128 * sub + lane_base + lane;      // this thread's feature index (MMA-M)
    const bool f_ok = f < ME;
    const GemmDesc g = op.g;             // by value: descriptor fields must not be re-read from HBM around every store
    const float debias = f16 ? 1.0f : TF32_TRUNC_DEBIAS;   // fp16 operand copies are rounded to nearest: nothing to remove
    const int epi = op.epi;
    float* const adamP = op.P; float* const adamM = op.Mo; float* const adamV = op.Vo;
    const int ncols = min(min(bn, NE - n0), cbeg + chalf);      // this warp's column range is [cbeg, ncols)
    const uint32_t trow = tmem_base + 256u * sub + ((uint32_t)lane_base << 16);
    uint32_t key0 = 0, key1 = 0, step = 0;
    float lr_t = 0.f;
    const bool noisy = (epi == EPI_FWD) && g.C2 != nullptr && g.sigma != 0.f;
    if (noisy || epi == EPI_ADAM) {
      const FoldState& fs = folds[g.fold];
      key0 = fs.key0; key1 = fs.key1; step = (uint32_t)fs.rng_step; lr_t = fs.lr_t[op.net];
    }

    if (epi == EPI_ADAM) {
      // Keras-2.0.9 Adam fused into the dW epilogue: the gradient never leaves the SM.  Accumulator D[n = f, k];
      // the weight tensor is [k, n] row-major, so for a fixed k the 32 lanes touch one 128-byte line of each of
      // W, m, v.  The three streams are register double-buffered one 16-column chunk ahead, and the first chunk
      // is requested BEFORE the accumulator is ready, so HBM latency overlaps the TMA/MMA phase.
      const float b1 = hp.b1, b2 = hp.b2, c1 = 1.0f - hp.b1, c2 = 1.0f - hp.b2, eps = hp.eps;
      const float ginv_a = (om.mode == 2) ? 1.0f / om.gscale : 1.0f;      // the dZ operand carries the loss scale
      // 8-column chunks: 2 x 24 prefetch registers keep the kernel under the 2-CTA/SM register budget (no spills,
      // which would force every load to be waited for immediately)
      auto fetch = [&](int c0, float (&pw)[8], float (&pm)[8], float (&pv)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (f_ok && c0 + j < ncols) {
            const size_t idx = (size_t)(n0 + c0 + j) * g.ldc + f;
            pw[j] = __ldcs(adamP + idx); pm[j] = __ldcs(adamM + idx); pv[j] = __ldcs(adamV + idx);
          }
        }
      };
      auto apply = [&](int c0, const float (&v)[8], const float (&pw)[8], const float (&pm)[8], const float (&pv)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (f_ok && c0 + j < ncols) {
            const size_t idx = (size_t)(n0 + c0 + j) * g.ldc + f;
            const float gr = v[j] * ginv_a;
            const float m = fmaf(b1, pm[j], c1 * gr), vv = fmaf(b2, pv[j], c2 * gr * gr);
            __stcs(adamM + idx, m); __stcs(adamV + idx, vv);
            const float w = pw[j] - lr_t * __fdividef(m, sqrtf(vv) + eps);
            __stcs(adamP + idx, w);
            if (om.mode == 2) om.hbase[adamP + idx - om.fbase] = __float2half_rn(w);
          }
        }
      };
      float pwA[8], pmA[8], pvA[8], pwB[8], pmB[8], pvB[8], v[8];
      fetch(cbeg, pwA, pmA, pvA);
      fetch(cbeg + 8, pwB, pmB, pvB);
      mbar_wait(tmem_full, 0);
      fence_after();
      for (int c0 = cbeg; c0 < ncols; c0 += 16) {
        tmem_ld8(trow + (uint32_t)c0, v);
        apply(c0, v, pwA, pmA, pvA);
        if (c0 + 16 < ncols) fetch(c0 + 16, pwA, pmA, pvA);
        if (c0 + 8 < ncols) {
          tmem_ld8(trow + (uint32_t)(c0 + 8), v);
          apply(c0 + 8, v, pwB, pmB, pvB);
          if (c0 + 24 < ncols) fetch(c0 + 24, pwB, pmB, pvB);
        }
      }
    } else if (epi == EPI_DX) {
      // dZ_prev[r, f] = acc * act'(h_prev[r, f]); h is prefetched one chunk ahead (and before the accumulator is ready)
      auto fetch = [&](int c0, float (&av)[16]) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          av[j] = (f_ok && g.act != ACT_NONE && c0 + j < ncols) ? __ldg(g.aux + (size_t)(n0 + c0 + j) * g.ldaux + f) : 1.0f;
      };
      auto apply = [&](int c0, const float (&v)[16], const float (&av)[16]) {
        if (!f_ok) return;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (c0 + j >= ncols) break;
          float x = v[j] * debias;
          if (g.act == ACT_RELU) x = (av[j] > 0.f) ? x : 0.f;
          else if (g.act == ACT_SOFTPLUS) x *= 1.0f - __expf(-av[j]);
          float* const pc = g.C + (size_t)(n0 + c0 + j) * g.ldc + f;
          if (om.mode == 2 && (g.rnd & 1)) put_grad16(om.hbase + (pc - om.fbase), x, om);    // operand only: 16-bit copy, loss-scaled like acc
          else *pc = (g.rnd & 1) ? rna_tf32(x) : x;
        }
      };
      // The h values do not depend on the accumulator either: whole 16-row chunks of them are loaded while the mainloop
      // runs and parked in the free TMEM columns (same split as the forward epilogue's noise); what does not fit is
      // register-prefetched one chunk ahead as before.
      float avA[16], avB[16], v[16];
      int cpk = cbeg;                                   // rows [cbeg, cpk) of this thread's range are parked
      uint32_t taux = 0;
      if (g.act != ACT_NONE) {
        int fbeg, flen;
        if (MT == 2) { fbeg = 256 * sub + bn; flen = 256 - bn; }
        else { flen = ((TMEM_COLS - bn) / (EPW == 8 ? 2 : 1)) & ~15; fbeg = bn + wg * flen; }
        const int nchunk = min(flen >> 4, (ncols - cbeg + 15) >> 4);
        taux = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)fbeg;
        if (nchunk > 0) {
          fetch(cbeg, avA);
          for (int k = 0; k < nchunk; k += 2) {
            if (k + 1 < nchunk) fetch(cbeg + 16 * (k + 1), avB);
            tmem_st16(taux + 16u * (uint32_t)k, avA);
            if (k + 1 < nchunk) {
              if (k + 2 < nchunk) fetch(cbeg + 16 * (k + 2), avA);
              tmem_st16(taux + 16u * (uint32_t)(k + 1), avB);
            }
          }
          tmem_wait_st();
          cpk = cbeg + 16 * nchunk;
        }
      }
      if (cpk < ncols) fetch(cpk, avA);
      mbar_wait(tmem_full, 0);
      fence_after();
      for (int c0 = cbeg; c0 < min(cpk, ncols); c0 += 16) {
        tmem_ld16(trow + (uint32_t)c0, v);
        tmem_ld16(taux + (uint32_t)(c0 - cbeg), avB);
        apply(c0, v, avB);
      }
      for (int c0 = cpk; c0 < ncols; c0 += 32) {
        tmem_ld16(trow + (uint32_t)c0, v);
        if (c0 + 16 < ncols) fetch(c0 + 16, avB);
        apply(c0, v, avA);
        if (c0 + 16 < ncols) {
          tmem_ld16(trow + (uint32_t)(c0 + 16), v);
          if (c0 + 32 < ncols) fetch(c0 + 32, avA);
          apply(c0 + 16, v, avB);
        }
      }
    } else {
      // GaussianNoise of the next layer's input (EPI_FWD with C2): the draws do not depend on the accumulator, and the
      // epilogue warps are idle while the TMA/MMA warps run the mainloop -- so they draw now and park the values in the
      // TMEM columns the accumulator leaves free (sub-tile s owns columns [256 s + bn, 256 s + 256); with one sub-tile the
      // two warp groups share [bn, TMEM_COLS)).  Row groups that do not fit are drawn in the epilogue as before.
      int npre = 0;
      uint32_t tnoise = 0;
      if (epi == EPI_FWD && noisy && ((g.row0 + n0 + cbeg) & 3) == 0 && (hp.dp_bg == hp.dp_bloc || (hp.dp_bloc & 3) == 0)) {
        int fbeg, flen;
        if (MT == 2) { fbeg = 256 * sub + bn; flen = 256 - bn; }
        else { flen = ((TMEM_COLS - bn) / (EPW == 8 ? 2 : 1)) & ~3; fbeg = bn + wg * flen; }
        npre = min(flen >> 2, (ncols - cbeg + 3) >> 2);
        tnoise = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)fbeg;
        for (int gi = 0; gi < npre; ++gi) {
          float nz[4];
          normal4(key0, key1, (uint32_t)global_row(g.row0 + n0 + cbeg + 4 * gi, hp) >> 2, (uint32_t)f, step, (uint32_t)g.tid, nz);
          tmem_st4(tnoise + 4u * (uint32_t)gi, nz);
        }
        if (npre > 0) tmem_wait_st();
      }
      mbar_wait(tmem_full, 0);
      fence_after();
      for (int c0 = cbeg; c0 < ncols; c0 += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c0, v);                   // warp-collective: every lane takes part
        if (epi == EPI_FWD) {
          // rows r = n0 + c0 + j; noise is grouped by 4 consecutive rows of one column (= this feature)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r0 = n0 + c0 + 4 * q;
            if (r0 >= NE) break;
            float nz[4] = {0.f, 0.f, 0.f, 0.f};
            if (noisy) {
              const int gi = ((c0 - cbeg) >> 2) + q;
              if (gi < npre)                                 // warp-uniform: the TMEM load stays convergent
                tmem_ld4(tnoise + 4u * (uint32_t)gi, nz);
              else if (((g.row0 + r0) & 3) == 0 && (hp.dp_bg == hp.dp_bloc || (hp.dp_bloc & 3) == 0))
                normal4(key0, key1, (uint32_t)global_row(g.row0 + r0, hp) >> 2, (uint32_t)f, step, (uint32_t)g.tid, nz);
              else
                for (int i = 0; i < 4; ++i) nz[i] = normal1(key0, key1, (uint32_t)global_row(g.row0 + r0 + i, hp), (uint32_t)f, step, (uint32_t)g.tid);
            }
            if (!f_ok) continue;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = r0 + i;
              if (r >= NE) break;
              float x = v[4 * q + i] * debias;
              if (g.act == ACT_RELU) x = fmaxf(x, 0.f);
              else if (g.act == ACT_SOFTPLUS) x = softplusf(x);
              if (g.C) {            // clean activation: kept in fp32 (act' of the backward pass, feature matching) ...
                float* const pc = g.C + (size_t)r * g.ldc + f;
                *pc = ((g.rnd & 1) && om.mode == 1) ? rna_tf32(x) : x;
                if (om.mode == 2 && (g.rnd & 1)) om.hbase[pc - om.fbase] = __float2half_rn(x);     // ... plus its operand copy
              }
              if (g.C2) {           // noisy activation: only ever a GEMM operand
                const float y = x + g.sigma * nz[i];
                float* const pc2 = g.C2 + (size_t)r * g.ldc2 + f;
                if (om.mode == 2 && (g.rnd & 2)) om.hbase[pc2 - om.fbase] = __float2half_rn(y);
                else *pc2 = (g.rnd & 2) ? rna_tf32(y) : y;
              }
            }
          }
        } else {   // EPI_STORE: dW into the flat gradient buffer (data-parallel mode: all-reduced before Adam)
          if (!f_ok) continue;
          float* const dst = ksplit > 1 ? op.ws + (size_t)kslice * op.ws_stride : g.C;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = n0 + c0 + j;
            if (k >= NE) break;
            dst[(size_t)k * g.ldc + f] = v[j];
          }
        }
      }
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
  }
}

// Sums the `ks` contraction slices of a split-K dW in slice order (bitwise reproducible) into the gradient tensor.
__global__ void __launch_bounds__(256) k_splitk_reduce(const TcOp* __restrict__ ops, int ks) {
  const TcOp& op = ops[blockIdx.y];
  const size_t n4 = (size_t)op.NE * op.g.ldc / 4, stride4 = (size_t)op.ws_stride / 4;
  const float4* __restrict__ ws = reinterpret_cast<const float4*>(op.ws);
  float4* __restrict__ out = reinterpret_cast<float4*>(op.g.C);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 a = ws[i];
    for (int s = 1; s < ks; ++s) {
      const float4 b = ws[(size_t)s * stride4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    out[i] = a;
  }
}

// ====================================================================================================
// k_dw_adam_tc -- dW (+db) on tcgen05 with the Keras-2.0.9 Adam update fused in, optimizer state staged by TMA.
//
// grad(Waug_l)[k, n] = sum_r A[r, k] dZ[r, n]  as  D[n (128 lanes), k (128 columns)], then for the same tile
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; W -= lr_t m / (sqrt(v) + eps)              
// W, m and v tiles never go through the LSU: one thread streams them HBM -> smem -> HBM in [KC rows x 128 cols]
// chunks with cp.async.bulk.tensor (full 128-byte lines, deep queues); the 4 epilogue warps combine each chunk
// with the matching 8 accumulator columns from TMEM in shared memory.  24 B per parameter of HBM traffic, which
// is the algorithmic minimum for Adam; the gradient never leaves the SM.
//   smem: 2 operand stages x 32 KB + NB chunk buffers x (3 x KC x 512 B); 2 CTAs per SM.
// ====================================================================================================
struct alignas(64) TcAdamOp {
  CUtensorMap mapA, mapB;          // dZ[r, n] and A[r, k], both MN-major operands (32-byte-atom 128B swizzle)
  CUtensorMap mapP, mapM, mapV;    // Waug, m, v as [rows k, cols n], box = 128 cols x KC rows, no swizzle
  int ME, NE, KE;                  // out-features n, in-features(+1) k, contraction rows r
  int fold;
  int net;
  int esz;                         // operand element size (4 = tf32, 2 = fp16 copies)
  int afmt;                        // esz == 2: format of the dZ operand (0 = f16, 1 = bf16); the activation operand is f16
  int ldh;                         // pitch of W and of its fp16 operand copy
  float ginv;                      // 1 / loss scale carried by the dZ operand (1 unless fp16)
  __half* Ph;                      // fp16 operand copy of W, refreshed with every update (null unless fp16)
};

#define TCA_KC 8
#define TCA_NB 4

namespace tc {
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace tc

__global__ void __launch_bounds__(192, 2)
k_dw_adam_tc(const TcAdamOp* __restrict__ ops, FoldState* __restrict__ folds, AdamHyper hp) {
  using namespace tc;
  pdl_launch_dependents();
  constexpr int STAGES = 2, KC = TCA_KC, NB = TCA_NB;
  constexpr uint32_t STAGE_BYTES = 2 * 128 * 128;            // A (4 x 4 KB blocks) + B (4 x 4 KB blocks)
  constexpr uint32_t ARR_BYTES = KC * 128 * 4;               // one array (W, m or v) of one chunk
  constexpr uint32_t CHUNK_BYTES = 3 * ARR_BYTES;
  // chunk buffers: NB dedicated ones + NX carved out of the operand stages once the MMAs have consumed them
  constexpr int NX = (STAGES * STAGE_BYTES) / CHUNK_BYTES;
  constexpr int NBUF = NB + NX;
  const TcAdamOp& op = ops[blockIdx.z];
  const int ME = op.ME, NE = op.NE, KE = op.KE;
  const int m0 = blockIdx.x * 128, n0 = blockIdx.y * 128;     // n-feature tile (lanes), k-feature tile (columns)
  if (m0 >= ME || n0 >= NE) return;

  extern __shared__ __align__(1024) uint8_t smem_dyn[];
  uint8_t* smem = smem_dyn;
  if ((smem_u32(smem) & 1023u) != 0) __trap();                // swizzled TMA tiles need 1024-byte alignment
  uint8_t* cbuf = smem + STAGES * STAGE_BYTES;                // NB dedicated chunk buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(cbuf + NB * CHUNK_BYTES);
  uint64_t* full = bars;                       // [STAGES]   operands landed
  uint64_t* empty = bars + STAGES;             // [STAGES]   operands consumed
  uint64_t* tmem_full = bars + 2 * STAGES;     //            accumulator complete (2 waiters: epilogue, DMA thread)
  uint64_t* cfull = bars + 2 * STAGES + 1;     // [NBUF]     optimizer-state chunk landed
  uint64_t* cdone = cfull + NBUF;              // [NBUF]     chunk updated in smem (128 arrivals)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(cdone + NBUF);
  auto buf_ptr = [&](int b) -> uint8_t* { return b < NB ? cbuf + (size_t)b * CHUNK_BYTES : smem + (size_t)(b - NB) * CHUNK_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool f16 = op.esz == 2;                 // see k_gemm_tc: 64 contraction rows per stage, 8 KB MN-major boxes
  const int kblk = f16 ? 64 : TC_KBLK;
  const int nkb = (KE + kblk - 1) / kblk;
  const int ncols = min(128, NE - n0);
  const int nch = (ncols + KC - 1) / KC;

  if (warp == 0 && lane == 0) {
    prefetch_map(&op.mapA); prefetch_map(&op.mapB); prefetch_map(&op.mapP); prefetch_map(&op.mapM); prefetch_map(&op.mapV);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    for (int b = 0; b < NBUF; ++b) { mbar_init(&cfull[b], 1); mbar_init(&cdone[b], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      // ---- DMA thread: optimizer-state chunks first (they do not depend on the MMA), then the operands ----
      auto load_chunk = [&](int c) {
        const int b = c % NBUF;
        uint8_t* dst = buf_ptr(b);
        mbar_expect_tx(&cfull[b], CHUNK_BYTES);
        tma_load_2d(&op.mapP, &cfull[b], dst, m0, n0 + c * KC);
        tma_load_2d(&op.mapM, &cfull[b], dst + ARR_BYTES, m0, n0 + c * KC);
        tma_load_2d(&op.mapV, &cfull[b], dst + 2 * ARR_BYTES, m0, n0 + c * KC);
      };
      for (int c = 0; c < NB && c < nch; ++c) load_chunk(c);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], STAGE_BYTES);
        uint8_t* sa = smem + (size_t)s * STAGE_BYTES;
        uint8_t* sb = sa + 128 * 128;
        const int k0 = kb * kblk;
        const int bbytes = f16 ? 8192 : 4096;
        for (int b = 0; b < 128 / kblk; ++b) tma_load_2d(&op.mapA, &full[s], sa + b * bbytes, m0 + kblk * b, k0);
        for (int b = 0; b < 128 / kblk; ++b) tma_load_2d(&op.mapB, &full[s], sb + b * bbytes, n0 + kblk * b, k0);
      }
      // the MMAs have consumed every operand stage once the accumulator is complete: the operand region now takes
      // NX more chunks in flight
      mbar_wait(tmem_full, 0);
      for (int c = NB; c < NBUF && c < nch; ++c) load_chunk(c);
      // ---- write-back of updated chunks; a buffer is refilled one chunk late so that the store which is reading it has
      //      had a whole chunk period to drain (wait_group.read 1 instead of a full stall per chunk) ----
      for (int c = 0; c < nch; ++c) {
        const int b = c % NBUF;
        mbar_wait(&cdone[b], (uint32_t)(c / NBUF) & 1u);
        const uint8_t* src = buf_ptr(b);
        tma_store_2d(&op.mapP, src, m0, n0 + c * KC);
        tma_store_2d(&op.mapM, src + ARR_BYTES, m0, n0 + c * KC);
        tma_store_2d(&op.mapV, src + 2 * ARR_BYTES, m0, n0 + c * KC);
        bulk_commit();
        if (c >= 1 && c - 1 + NBUF < nch) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          load_chunk(c - 1 + NBUF);
        }
      }
      bulk_wait_read0();
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t fmtA = f16 ? (uint32_t)op.afmt : 2u, fmtB = f16 ? 0u : 2u;
      const uint32_t idesc = (1u << 4) | (fmtA << 7) | (fmtB << 10) | (1u << 15) | (1u << 16) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t lbo = f16 ? 8192u : 4096u, sbo = f16 ? 1024u : 512u, lay = f16 ? 2u : 1u, kstep = f16 ? 2048u : 1024u;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (uint32_t)(kb / STAGES) & 1u;
        mbar_wait(&full[s], ph);
        fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * STAGE_BYTES), sb = sa + 128 * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = smem_desc(sa + k * kstep, lbo, sbo, lay), db = smem_desc(sb + k * kstep, lbo, sbo, lay);
          if (f16) mma_f16(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          else mma_tf32(tmem_base, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
        }
        mma_commit(&empty[s]);
      }
      mma_commit(tmem_full);
    }
  } else {
    // ---- epilogue: 4 warps, thread = output feature n (TMEM lane = smem column) ----
    const int lane_base = 32 * (warp & 3);
    const int nl = lane_base + lane;
    const uint32_t trow = tmem_base + ((uint32_t)lane_base << 16);
    const float lr_t = folds[op.fold].lr_t[op.net];
    const float b1 = hp.b1, b2 = hp.b2, c1 = 1.0f - hp.b1, c2 = 1.0f - hp.b2, eps = hp.eps;
    const float ginv = op.ginv;
    __half* const Ph = op.Ph;
    const int ldh = op.ldh;
    const bool n_ok = m0 + nl < ME;
    mbar_wait(tmem_full, 0);
    fence_after();
    for (int c = 0; c < nch; ++c) {
      const int b = c % NBUF;
      float* sP = reinterpret_cast<float*>(buf_ptr(b));
      float* sM = sP + KC * 128;
      float* sV = sM + KC * 128;
      float g[KC];
      tmem_ld8(trow + (uint32_t)(c * KC), g);
      mbar_wait(&cfull[b], (uint32_t)(c / NBUF) & 1u);
#pragma unroll
      for (int j = 0; j < KC; ++j) {
        const int i = j * 128 + nl;
        const float gr = g[j] * ginv;
        const float m = fmaf(b1, sM[i], c1 * gr), v = fmaf(b2, sV[i], c2 * gr * gr);
        sM[i] = m; sV[i] = v;
        const float w = sP[i] - lr_t * __fdividef(m, sqrtf(v) + eps);
        sP[i] = w;
        // fp16 operand copy of the updated weights for the next forward / dX (2 more bytes per parameter; a warp writes
        // 64 contiguous bytes per row).  The TMA store of the fp32 tile clips at the tensor's extents; this store must too.
        if (Ph && n_ok && c * KC + j < ncols) Ph[(size_t)(n0 + c * KC + j) * ldh + m0 + nl] = __float2half_rn(w);
      }
      fence_proxy_async();              // generic-proxy writes -> visible to the bulk store
      mbar_arrive(&cdone[b]);
    }
  }

  fence_before();
  __syncthreads();
  if (warp == 1) {
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}
