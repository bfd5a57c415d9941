// Tap-gather GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
//   out[m,:] = sum_tap mean_{j in nbr(m,tap)} [A[j,:] | onehot(type_j)] . W[tap]  (+bias +row_add +resid)
//
// This is the B200 replacement of the reference's GraphConv op sequence
//   x[col] (aten::index) -> scatter_mean into a [7N, C] buffer -> view(N, 7C) @ W
// (reference models/networks/modules.py:194-220, diffusion_networks/utils/scatter.py:42-66) and of
// its dense Conv3d / Conv1x1 / Down/Upsample GEMMs (modules.py:332-339, 392-395, 440-443, 493-502).
// The im2col buffer is never written to HBM: gather warps build each [128 x 64] bf16 A tile
// directly in shared memory, in the 128-byte-swizzled K-major layout tcgen05.mma reads.
//
// One persistent CTA per SM, 448 threads, warp-specialised:
//   warps 0-3   epilogue: tcgen05.ld of the fp32 accumulator, +bias/+emb[batch]/+residual, store
//   warp  4     MMA issuer (whole warp walks the pipeline, elect.sync lane issues): tcgen05.mma 128 x BN x 16,
//               tcgen05.commit releases the shared-memory stages
//   warp  5     weight loader: cp.async.bulk (TMA 1-D) of pre-swizzled [BN x 64] tiles
//   warps 6-13  gather producers (4 groups x 2 warps, each group owns every 4th K block):
//               tap table -> sixteen 16-byte cp.async (LDGSTS) per thread straight into the swizzled A stage,
//               completion by cp.async.mbarrier.arrive.noinc; no registers, no waiting.  The producers' address
//               arithmetic is kept to 4 instructions per copy (it shares issue slots with the MMA warp).
// Pipelines: two shared-memory rings (A: gathered tiles, deep; B: weight tiles, shallow) with full/empty
// mbarriers between {producers, loader} and the MMA warp; a stage holds KSUB = 2 K blocks when BN <= 128 so that
// the MMA warp's per-stage overhead is amortised over 8 MMAs; two TMEM accumulators (full/empty mbarriers)
// between MMA and epilogue, so tile i+1 is computed while tile i drains.
//
// K is consumed as 64-wide blocks ordered (channel block outer, tap inner): the 7 (or 27) taps
// of one 64-channel slab touch the same few hundred source rows, which then sit in L2.
// The one-hot node-type columns (modules.py:199-202) are one extra K block of per-slot type fractions, read
// from a per-graph precomputed tensor (of_graph_type_block); slots with several finer neighbours read a
// pre-averaged row (of_gather_mean_rows); the weights are re-laid once by of_pack_weight_tc.
// Measurements behind each of these choices: profiles/tc_gather_experiments_r01.md.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>
#include <cuda.h>          // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)

namespace of {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                 // bf16 elements = 128 bytes = one swizzle row
constexpr int TC_EPI_WARPS = 4;
constexpr int TC_PROD_WARPS = 8;
constexpr int TC_THREADS = (TC_EPI_WARPS + 2 + TC_PROD_WARPS) * 32;   // 448
constexpr int TC_MAX_TAPS = 27;
constexpr int TC_GROUPS = 4;                // producer groups of 2 warps

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a trap (an error the host sees), never a hung GPU.
// variant for the long waits of the epilogue warps: back off between polls so that the four idle warps do not
// compete with the producers for issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(200);
    if (clock64() - t0 > 4000000000ll) {
      printf("octfusion_b200 gemm_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      printf("octfusion_b200 gemm_tc: mbarrier timeout (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
// one lane of a fully converged warp (elect.sync): the uniform-datapath instructions (UTCHMMA, UTCBAR, UBLKCP)
// are then issued from converged code instead of the ELECT/BRA.U.ANY retry loops the compiler emits for a
// divergent `if (lane == 0)` region
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// TMA gather4: four arbitrary rows x 64 columns (4 x 128 B) of a 2-D tensor -> 512 contiguous bytes of shared
// memory, 128B-swizzled by the hardware; rows outside the tensor are zero-filled.
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tmap, int col, int r0, int r1, int r2, int r3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t result_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(result_slot), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, M=128, N=BN, K=16
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same, descriptors given by their low words; high word = SBO 1024 B (>>4 = 64) | version 1 @46 | SWIZZLE_128B (2) @61
constexpr uint32_t TC_DESC_HI = 64u | (1u << 14) | (2u << 29);
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(TC_DESC_HI)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants: EXPERIMENTAL, see the note above gather_gemm_tc_kernel ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t result_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(result_slot), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit2(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
// D[tmem of both CTAs] (+)= A (128 rows from each CTA) * B (BN/2 rows from each CTA): M = 256, N = BN, K = 16
__device__ __forceinline__ void umma_bf16_lo2(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(TC_DESC_HI)
      : "memory");
}
// K-major, SWIZZLE_128B operand descriptor (cute::UMMA::SmemDescriptor): start>>4 in [0,14),
// LBO>>4 in [16,30) (unused for swizzled K-major: 1), SBO>>4 in [32,46) = 1024 B between 8-row
// groups, version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32=1 @4, a_format BF16=1 @7, b_format BF16=1 @10,
// a/b K-major (0) @15/@16, N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t make_idesc_pair(int bn) {          // cta_group::2: M = 256
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc(int bn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

#define OF_TMEM_LD32(taddr, r)                                                                               \
  asm volatile(                                                                                              \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                              \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,"  \
      "%27,%28,%29,%30,%31}, [%32];"                                                                         \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), \
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),          \
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),          \
        "=r"(r[30]), "=r"(r[31])                                                                             \
      : "r"(taddr)                                                                                           \
      : "memory")
#define OF_TMEM_LD16(taddr, r)                                                                               \
  asm volatile(                                                                                              \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                              \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                                      \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) \
      : "r"(taddr)                                                                                           \
      : "memory")
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 16-byte asynchronous global->shared copy (LDGSTS); src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------
// shared-memory plan
// ------------------------------------------------------------------------------------------------
template <int BN, int CG = 1>
struct TcCfg {
  // One pipeline stage holds KSUB consecutive 64-wide K blocks: the MMA warp then spends its ~0.25 us of waits,
  // election, descriptor set-up and commits once per KSUB*4 MMAs.  For BN <= 128 four MMAs (<= 256 cycles of tensor
  // work) are shorter than that loop overhead and the tensor pipe starved (profiles/tc_gather_experiments_r01.md).
  static constexpr int KSUB = BN <= 128 ? 2 : 1;
  static constexpr int A_SUB_BYTES = TC_BM * 128;                   // 16 KB: one K block of A
  static constexpr int B_SUB_BYTES = (BN / CG) * 128;               // one K block of B (a CTA pair stages half each)
  static constexpr int A_BYTES = KSUB * A_SUB_BYTES;
  static constexpr int B_BYTES = KSUB * B_SUB_BYTES;
  // Two independent rings.  The gathered A tiles need depth: their throughput is (bytes in flight) / (~0.7 us), see
  // profiles/tc_gather_experiments_r01.md.  The weight tiles stream from L2 by TMA and need only a shallow ring.
  static constexpr int B_STAGES = KSUB > 1 ? 2 : (BN >= 64 ? 3 : 4);
  static constexpr int A_STAGES = (192 * 1024 - B_STAGES * B_BYTES) / A_BYTES;      // 6 x 16 KB (BN=256), 4-5 x 32 KB
  static constexpr int STAGES = A_STAGES;                                           // (A ring depth)
  static constexpr int TMEM_COLS = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128 : 2 * BN <= 256 ? 256 : 512;
  static constexpr int TAP_BYTES = 0;                               // (tap table is read through L1)
  static constexpr int AUX_BYTES = CG == 2 ? 1024 : 512;            // mbarriers + tmem slot (+ the pair's peer-ready ring)
  static constexpr int RING_BYTES = A_STAGES * A_BYTES + B_STAGES * B_BYTES;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + RING_BYTES + TAP_BYTES + AUX_BYTES;
};

struct TcParams {
  of_gemm_args g;
  int num_kb;        // K blocks per tile
  int cblocks;       // (c0+c1)/64
  int npad;          // N rounded up to 16 (rows per K block in the packed weight image)
  int m_tiles, n_tiles;
  int use_tma;       // 1: half of the row groups of every feature K block are fetched by TMA gather4
  int rows0, rows1;  // row counts of a0 / a1 (TMA out-of-bounds row = zero fill for empty slots)
  int debug;         // OCTFUSION_TC_DEBUG bit mask (timing experiments only): 1 no gather, 2 no weight copy, 4 no epilogue I/O, 8 no MMA, 32 no tap-table reads, 64 no weight ring at all
};

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) gather_gemm_tc_kernel(const TcParams p,
                                                                       const __grid_constant__ CUtensorMap tmap0,
                                                                       const __grid_constant__ CUtensorMap tmap1) {
  using Cfg = TcCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stage_base = smem_base;
  const uint32_t b_ring = smem_base + Cfg::A_STAGES * Cfg::A_BYTES;
  const uint32_t aux = smem_base + Cfg::RING_BYTES + Cfg::TAP_BYTES;
  // aux layout: a_full[16] | a_empty[16] | b_full[4] | b_empty[4] | tmem_full[2] | tmem_empty[2] | tmem slot
  const uint32_t bar_full = aux, bar_empty = aux + 128;
  const uint32_t bar_bfull = aux + 256, bar_bempty = aux + 288;
  const uint32_t bar_tfull = aux + 320, bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::RING_BYTES + Cfg::TAP_BYTES + 352);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const of_gemm_args& g = p.g;
  const int taps = g.taps;
  const int total_tiles = p.m_tiles * p.n_tiles;

  if (warp == TC_EPI_WARPS && lane == 0) {
    for (int s = 0; s < Cfg::A_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, Cfg::KSUB * (TC_PROD_WARPS / TC_GROUPS) * 32);   // every producer thread of the stage's K blocks
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < Cfg::B_STAGES; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, TC_EPI_WARPS * 32);
    }
    fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp < TC_EPI_WARPS) {
    // =========================== epilogue ===========================
    const int r = warp * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int ptile = g.reverse ? total_tiles - 1 - tile : tile;
      const int m0 = (ptile / p.n_tiles) * TC_BM, n0 = (ptile % p.n_tiles) * BN;
      const int as = it & 1;
      if (p.debug & 8192) mbar_wait(bar_tfull + 8 * as, (it >> 1) & 1);
      else mbar_wait_relaxed(bar_tfull + 8 * as, (it >> 1) & 1);
      tc_fence_after();
      const int m = m0 + r;
      const bool row_ok = m < g.M;
      const int64_t orow = row_ok ? (g.out_rows ? (int64_t)g.out_rows[m] : (int64_t)m) : 0;
      const float* radd = (row_ok && g.row_add) ? g.row_add + (int64_t)g.row_add_idx[m] * g.ld_row_add : nullptr;
      const __nv_bfloat16* res =
          (row_ok && g.resid) ? reinterpret_cast<const __nv_bfloat16*>(g.resid) + (int64_t)m * g.ld_resid : nullptr;
      constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t acc[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * BN + c0);
        if (CH == 32) { OF_TMEM_LD32(taddr, acc); } else { OF_TMEM_LD16(taddr, acc); }
        tmem_ld_wait();
        if (!row_ok || (p.debug & 4)) continue;
        const int nb = n0 + c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(acc[j]);
        const bool full = (nb + CH <= g.N);
        if (g.bias) {
#pragma unroll
          for (int j = 0; j < CH; ++j) if (full || nb + j < g.N) v[j] += g.bias[nb + j];
        }
        if (radd) {
#pragma unroll
          for (int j = 0; j < CH; ++j) if (full || nb + j < g.N) v[j] += radd[nb + j];
        }
        if (res) {
          if (full && (g.ld_resid % 8 == 0)) {
#pragma unroll
            for (int q = 0; q < CH / 8; ++q) {
              float f[8];
              bf16x8_to_f32(ldg_nc_v4(res + nb + q * 8), f);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[q * 8 + j] += f[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) if (nb + j < g.N) v[j] += __bfloat162float(res[nb + j]);
          }
        }
        if (g.out_f32) {
          float* o = reinterpret_cast<float*>(g.out) + orow * g.ldo + nb;
          if (full && (g.ldo % 4 == 0)) {
#pragma unroll
            for (int q = 0; q < CH / 4; ++q)
              *reinterpret_cast<float4*>(o + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) if (nb + j < g.N) o[j] = v[j];
          }
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(g.out) + orow * g.ldo + nb;
          if (full && (g.ldo % 8 == 0)) {
#pragma unroll
            for (int q = 0; q < CH / 8; ++q) *reinterpret_cast<uint4*>(o + q * 8) = f32_to_bf16x8(v + q * 8);
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) if (nb + j < g.N) o[j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_tempty + 8 * as);
    }
  } else if (warp == TC_EPI_WARPS) {
    // =========================== MMA issuer ===========================
    // the whole warp walks the pipeline (all lanes wait on the barriers); one elected lane issues
    {
      constexpr uint32_t idesc = make_idesc(BN);
      int stage = 0, bstage = 0;
      uint32_t phase = 0, bphase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        mbar_wait(bar_tempty + 8 * as, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < p.num_kb; kb += Cfg::KSUB) {
          if (!(p.debug & 64)) mbar_wait(bar_bfull + 8 * bstage, bphase);
          mbar_wait(bar_full + 8 * stage, phase);
          if (p.debug & 256) tc_fence_after();     // (experiment) not needed: the operands arrive by cp.async / TMA, not tcgen05
          const uint32_t a_addr = stage_base + stage * Cfg::A_BYTES;
          const uint32_t b_addr = b_ring + bstage * Cfg::B_BYTES;
          if (elect_one()) {
            // descriptor low words: start address >> 4 (+2 per 32-byte K step), LBO = 1; the high word is constant
            const uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo = ((b_addr & 0x3FFFFu) >> 4) | (1u << 16);
            if (!(p.debug & 8)) {
#pragma unroll
              for (int j = 0; j < Cfg::KSUB; ++j) {
                if (j > 0 && kb + j >= p.num_kb) break;              // odd K-block count: the last stage is half full
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k)
                  umma_bf16_lo(d_tmem, a_lo + j * (Cfg::A_SUB_BYTES >> 4) + 2 * k, b_lo + j * (Cfg::B_SUB_BYTES >> 4) + 2 * k,
                               idesc, (kb > 0 || j > 0 || k > 0) ? 1u : 0u);
              }
            }
            umma_commit(bar_empty + 8 * stage);            // frees the A stage when these MMAs retire
            if (!(p.debug & 64)) umma_commit(bar_bempty + 8 * bstage);          // ... and the B stage
            if (kb + Cfg::KSUB >= p.num_kb) umma_commit(bar_tfull + 8 * as);   // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == Cfg::A_STAGES) { stage = 0; phase ^= 1; }
          if (++bstage == Cfg::B_STAGES) { bstage = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // =========================== weight loader ===========================
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint8_t* wp = reinterpret_cast<const uint8_t*>(g.w);
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n0 = ((g.reverse ? total_tiles - 1 - tile : tile) % p.n_tiles) * BN;
        for (int kb = 0; kb < p.num_kb && !(p.debug & 64); kb += Cfg::KSUB) {
          mbar_wait(bar_bempty + 8 * stage, phase ^ 1);
          const uint32_t b_addr = b_ring + stage * Cfg::B_BYTES;
          if (elect_one()) {
            if (p.debug & 2) { mbar_arrive(bar_bfull + 8 * stage); }
            else {
              const int nk = min(Cfg::KSUB, p.num_kb - kb);
              mbar_arrive_expect_tx(bar_bfull + 8 * stage, (uint32_t)nk * Cfg::B_SUB_BYTES);
              for (int j = 0; j < nk; ++j)
                bulk_g2s(b_addr + j * Cfg::B_SUB_BYTES, wp + ((int64_t)(kb + j) * p.npad + n0) * 128, Cfg::B_SUB_BYTES,
                         bar_bfull + 8 * stage);
            }
          }
          __syncwarp();
          if (++stage == Cfg::B_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =========================== gather producers ===========================
    // 4 independent groups of 2 warps; group g produces the K blocks whose running index is = g mod 4,
    // so 4 K blocks (64 KB of gathers) are in flight per SM and the memory latency of one block is
    // hidden behind the other three.  Each thread owns one 16-byte chunk column (q) of 16 rows.
    const int pt = threadIdx.x - (TC_EPI_WARPS + 2) * 32;           // 0..255
    const int grp = pt >> 6;                                        // producer group 0..3
    const int gt = pt & 63;
    const int q = gt & 7;                                           // 16-byte chunk of the 128-byte row
    const int rbase = gt >> 3;                                      // 0..7, rows rbase + 8*i
    const __nv_bfloat16* a0 = reinterpret_cast<const __nv_bfloat16*>(g.a0);
    const __nv_bfloat16* a1 = reinterpret_cast<const __nv_bfloat16*>(g.a1);
    const int32_t* __restrict__ tab = g.tap_tab;
    // K blocks of this CTA in consumption order: kbg = tile_iter * num_kb + kb; this group owns kbg = grp (mod 4).
    // The 16 table entries of the NEXT owned block are fetched before waiting for the current stage to be
    // released, which takes the table latency off the stage turnaround.
    const int my_tiles = (total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    // K-block slots of a tile, padded to whole stages (a pad slot carries no data: its owner only arrives)
    const uint32_t slots = (uint32_t)((p.num_kb + Cfg::KSUB - 1) / Cfg::KSUB * Cfg::KSUB);
    const uint32_t kb_total = (uint32_t)my_tiles * slots;
    const int feat_kb = p.cblocks * taps;
    // position of a K-block slot inside this CTA's work: (tile iteration, K block, channel block, tap), advanced
    // incrementally -- no integer divisions in the producer loop (its instruction stream competes with the MMA warp)
    struct Pos { int ti, kb, cb, tap; };
    auto norm = [&](Pos& s) {
      while (s.kb >= (int)slots) { s.kb -= (int)slots; ++s.ti; s.cb = 0; s.tap = s.kb; }
      while (s.tap >= taps) { s.tap -= taps; ++s.cb; }
    };
    auto tile_m0 = [&](int ti) {
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      const int pt = g.reverse ? total_tiles - 1 - tile : tile;
      return (p.n_tiles == 1 ? pt : pt / p.n_tiles) * TC_BM;
    };
    auto fetch_taps = [&](const Pos& s, int32_t* t) {
      const int kb = s.kb;
      const int m0 = tile_m0(s.ti);
      if (kb >= feat_kb || (p.debug & 32)) return;
      const int tap = s.tap;
      if (tab != nullptr) {
        const uint32_t base = (uint32_t)(m0 + rbase) * (uint32_t)taps + (uint32_t)tap;     // < 2^31 (checked on host)
#pragma unroll
        for (int i = 0; i < TC_BM / 8; ++i) {
          const int m = m0 + rbase + 8 * i;
          t[i] = -1;
          if (m < g.M) t[i] = __ldg(tab + (base + (uint32_t)(8 * i) * (uint32_t)taps));
        }
      } else if (g.in_rows != nullptr) {
#pragma unroll
        for (int i = 0; i < TC_BM / 8; ++i) {
          const int m = m0 + rbase + 8 * i;
          t[i] = -1;
          if (m < g.M) t[i] = __ldg(g.in_rows + m);
        }
      } else {
#pragma unroll
        for (int i = 0; i < TC_BM / 8; ++i) {
          const int m = m0 + rbase + 8 * i;
          t[i] = m < g.M ? m : -1;
        }
      }
    };
    int32_t tnext[TC_BM / 8];
    Pos cur{0, grp, 0, grp};
    norm(cur);
    if ((uint32_t)grp < kb_total) fetch_taps(cur, tnext);
    for (uint32_t kbg = (uint32_t)grp; kbg < kb_total; kbg += TC_GROUPS) {
      {
        const int kb = cur.kb;
        const int cur_cb = cur.cb;
        const int m0 = tile_m0(cur.ti);
        int32_t t[TC_BM / 8];
#pragma unroll
        for (int i = 0; i < TC_BM / 8; ++i) t[i] = tnext[i];
        cur.kb += TC_GROUPS; cur.tap += TC_GROUPS;
        norm(cur);                                           // now the position of kbg + TC_GROUPS
        if (kbg + TC_GROUPS < kb_total) fetch_taps(cur, tnext);
        const uint32_t sg = kbg / Cfg::KSUB;                       // stage counter; this K block is its sub-tile kbg % KSUB
        const uint32_t stage = sg % Cfg::A_STAGES;
        const uint32_t phase = (sg / Cfg::A_STAGES) & 1u;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t a_addr = stage_base + stage * Cfg::A_BYTES + (kbg % Cfg::KSUB) * Cfg::A_SUB_BYTES;
        if ((p.debug & 1) || kb >= p.num_kb) {
        } else if (kb < p.cblocks * taps) {
          const int cb = cur_cb;
          const int ch = cb * TC_BK;
          const __nv_bfloat16* src;
          int64_t ld;
          if (ch < g.c0) { src = a0 + ch; ld = g.lda0; } else { src = a1 + (ch - g.c0); ld = g.lda1; }
          src += q * 8;
          if (p.debug & 16) {
#pragma unroll
            for (int i = 0; i < TC_BM / 8; ++i) t[i] = 0;          // timing experiment: pure L1 hits
          }
          const __nv_bfloat16* msrc = reinterpret_cast<const __nv_bfloat16*>(g.a_multi) + ch + q * 8;
          // Two independent data paths share the gather.  Rows 8i+0..3 (i = 0..15) of the tile: TMA gather4 --
          // one instruction moves four source rows x 128 B and swizzles in hardware; the four row indices are
          // held by lanes q, q+8, q+16, q+24 of the group's first warp.  Rows 8i+4..7: 16-byte cp.async
          // (LDGSTS) through L1.  Either path alone saturates at ~16 KB/us per SM (request queues), together
          // they overlap.
          const bool tma_rows = p.use_tma && rbase < 4;           // this thread's rows belong to the TMA half
          if (p.use_tma && gt < 32) {                             // first warp of the group: all 32 lanes converge
            const CUtensorMap* tm = ch < g.c0 ? &tmap0 : &tmap1;
            const int col = ch < g.c0 ? ch : ch - g.c0;
            const int oob = ch < g.c0 ? p.rows0 : p.rows1;
#pragma unroll
            for (int i = 0; i < TC_BM / 8; ++i) {
              const int t1 = __shfl_sync(0xffffffffu, t[i], q + 8);
              const int t2 = __shfl_sync(0xffffffffu, t[i], q + 16);
              const int t3 = __shfl_sync(0xffffffffu, t[i], q + 24);
              if (rbase == 0 && (i & 7) == q) {                   // lane q issues groups i = q and q + 8
                const int t0 = t[i];
                const uint32_t dst = a_addr + (8 * i) * 128;
                if (t0 >= -1 && t1 >= -1 && t2 >= -1 && t3 >= -1) {
                  mbar_expect_tx(bar_full + 8 * stage, 512u);
                  tma_gather4(dst, tm, col, t0 < 0 ? oob : t0, t1 < 0 ? oob : t1, t2 < 0 ? oob : t2, t3 < 0 ? oob : t3,
                              bar_full + 8 * stage);
                } else {
                  // a multi-neighbour slot in the group: its pre-averaged row lives in another tensor -> LDGSTS
                  const int tt[4] = {t0, t1, t2, t3};
                  const __nv_bfloat16* s0 = src - q * 8;
                  const __nv_bfloat16* m0p = msrc - q * 8;
#pragma unroll
                  for (int r4 = 0; r4 < 4; ++r4) {
                    const int rr = 8 * i + r4;
                    const int tv = tt[r4];
                    const __nv_bfloat16* base = tv >= 0 ? s0 + (int64_t)tv * ld
                                                        : (tv == -1 ? s0 : m0p + (int64_t)(-(tv + 2)) * g.ld_multi);
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8)
                      cp_async_16(a_addr + rr * 128 + ((c8 ^ (rr & 7)) << 4), base + c8 * 8, tv == -1 ? 0u : 16u);
                  }
                }
              }
            }
          }
          if (!tma_rows) {
            // 16 asynchronous 16-byte global->shared copies back to back (no registers, no waiting):
            // one neighbour -> its row; none -> zero fill; several -> the pre-averaged row of a_multi.
            // Branch-free address: one select of (base, stride) + one 32x32+64 multiply-add per copy -- the producers'
            // instruction stream is what the MMA warp competes with for issue slots.
            const uint32_t dst0 = a_addr + rbase * 128 + ((q ^ (rbase & 7)) << 4);    // (rbase + 8i) & 7 == rbase & 7
            const uint64_t sbase = reinterpret_cast<uint64_t>(src), mbase = reinterpret_cast<uint64_t>(msrc);
            const uint32_t ldb = (uint32_t)ld * 2u, ldmb = (uint32_t)g.ld_multi * 2u;  // row strides in bytes
            // Multi-neighbour slots only occur on rows of coarse leaves (the first rows of the graph): two thirds of
            // the tiles have none, and then every copy is max / multiply-add / compare / LDGSTS.
            int32_t lo = t[0];
#pragma unroll
            for (int i = 1; i < TC_BM / 8; ++i) lo = min(lo, t[i]);
            if (!__any_sync(0xffffffffu, lo < -1)) {
#pragma unroll
              for (int i = 0; i < TC_BM / 8; ++i) {
                const int32_t tv = t[i];
                const uint64_t addr = sbase + (uint64_t)(uint32_t)max(tv, 0) * (uint64_t)ldb;
                cp_async_16(dst0 + i * 1024, reinterpret_cast<const void*>(addr), tv == -1 ? 0u : 16u);
              }
            } else {
#pragma unroll
              for (int i = 0; i < TC_BM / 8; ++i) {
                const int32_t tv = t[i];
                const bool multi = tv < -1;
                const uint32_t idx = multi ? (uint32_t)(-2 - tv) : (uint32_t)(tv < 0 ? 0 : tv);
                const uint64_t addr = (multi ? mbase : sbase) + (uint64_t)idx * (uint64_t)(multi ? ldmb : ldb);
                cp_async_16(dst0 + i * 1024, reinterpret_cast<const void*>(addr), tv == -1 ? 0u : 16u);
              }
            }
          }
        } else if (g.nt_block != nullptr) {
          // node-type block, precomputed per graph (of_graph_type_block): a plain coalesced copy of rows m0..m0+127
          const __nv_bfloat16* nb = reinterpret_cast<const __nv_bfloat16*>(g.nt_block) + q * 8;
#pragma unroll
          for (int i = 0; i < TC_BM / 8; ++i) {
            const int rr = rbase + 8 * i;
            const int m = m0 + rr;
            cp_async_16(a_addr + rr * 128 + ((q ^ (rr & 7)) << 4), m < g.M ? (const void*)(nb + (int64_t)m * 64) : (const void*)nb,
                        m < g.M ? 16u : 0u);
          }
        } else {
          // node-type block: column tap*ntype + type holds (#neighbours of that type)/(#neighbours)
          // = mean of the one-hot columns the reference concatenates (modules.py:199-202).
#pragma unroll 1
          for (int rr = gt; rr < TC_BM; rr += 64) {
            const uint32_t rowaddr = a_addr + rr * 128;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int c = 0; c < 8; ++c) sts_v4(rowaddr + (c << 4), z);
            const int m = m0 + rr;
            if (m >= g.M) continue;
            for (int tap = 0; tap < taps; ++tap) {
              const int32_t tv = tab ? __ldg(tab + (int64_t)m * taps + tap) : m;
              if (tv == -1) continue;
              unsigned long long packed;
              int n = 1;
              if (tv >= 0) {
                packed = 1ull << (8 * g.node_type[tv]);
              } else {
                packed = g.multi_types[-(tv + 2)];           // per-type neighbour counts of the slot
                n = 0;
                for (int ty = 0; ty < 8; ++ty) n += (int)((packed >> (8 * ty)) & 255ull);
              }
              for (int ty = 0; ty < g.ntype && ty < 8; ++ty) {
                const int c = (int)((packed >> (8 * ty)) & 255ull);
                if (c == 0) continue;
                const int col = tap * g.ntype + ty;
                const __nv_bfloat16 hv = __float2bfloat16_rn((float)c / (float)n);
                sts_u16(rowaddr + ((((col >> 3) ^ (rr & 7)) << 4) | ((col & 7) << 1)), __bfloat16_as_ushort(hv));
              }
            }
          }
          fence_proxy_async_smem();               // generic-proxy stores -> visible to the tensor core
        }
        // CUTLASS sm100 cp.async+UMMA protocol: one arrive that fires when this thread's cp.asyncs have
        // landed (self-incrementing, not counted) + one ordinary release-arrive (counted)
        // one counted arrival per thread: for gathered blocks it fires when this thread's cp.asyncs have landed
        // (cp.async.mbarrier.arrive.noinc); for the node-type block (generic stores + proxy fence) a plain arrive
        if ((kb < p.cblocks * taps || g.nt_block != nullptr) && kb < p.num_kb && !(p.debug & 1)) cp_async_mbar_arrive_noinc(bar_full + 8 * stage);
        else mbar_arrive(bar_full + 8 * stage);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_EPI_WARPS) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant of the kernel above (EXPERIMENTAL -- written in round 1, NOT YET RUN ON HARDWARE; opt-in with
// OCTFUSION_TC_CTA2=1, never dispatched otherwise, not covered by tests).  A copy of gather_gemm_tc_kernel with CG = 2
// (the `CG == 1` branches are dead here): two CTAs of a cluster compute a 256-row tile with
// tcgen05.mma.cta_group::2, each staging its own 128 rows of A and half of the B tile, which halves the weight bytes
// every SM writes to and re-reads from shared memory (profiles/tc_gather_experiments_r01.md, "Plan for the next step").
// Only rank 0 issues MMAs; rank 1's MMA warp relays "my A and B stages are full" to rank 0's peer-ready barriers;
// commits are multicast to both CTAs; rank 1's epilogue releases the accumulator on rank 0's barrier.
// Kept as a separate kernel so that the production kernel's code (and SASS) is untouched until this one is validated.
// ------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) gather_gemm_tc_pair_kernel(const TcParams p,
                                                                            const __grid_constant__ CUtensorMap tmap0,
                                                                            const __grid_constant__ CUtensorMap tmap1) {
  constexpr int CG = 2;
  using Cfg = TcCfg<BN, CG>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stage_base = smem_base;
  const uint32_t b_ring = smem_base + Cfg::A_STAGES * Cfg::A_BYTES;
  const uint32_t aux = smem_base + Cfg::RING_BYTES + Cfg::TAP_BYTES;
  // aux layout: a_full[16] | a_empty[16] | b_full[4] | b_empty[4] | tmem_full[2] | tmem_empty[2] | tmem slot
  const uint32_t bar_full = aux, bar_empty = aux + 128;
  const uint32_t bar_bfull = aux + 256, bar_bempty = aux + 288;
  const uint32_t bar_tfull = aux + 320, bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::RING_BYTES + Cfg::TAP_BYTES + 352);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const of_gemm_args& g = p.g;
  const int taps = g.taps;
  const int total_tiles = p.m_tiles * p.n_tiles;
  // work items of this CTA: tiles (CG = 1) or pair-tiles of 256 rows (CG = 2: both CTAs of a pair walk the same list)
  uint32_t cta_rank = 0;
  if constexpr (CG == 2) cta_rank = cluster_ctarank();
  const int w_first = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int w_stride = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int w_total = CG == 2 ? ((p.m_tiles + 1) / 2) * p.n_tiles : total_tiles;
  const uint32_t bar_pfull = aux + 512;                  // CG = 2: peer-ready ring (one barrier per A stage), leader's copy is used

  if (warp == TC_EPI_WARPS && lane == 0) {
    for (int s = 0; s < Cfg::A_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, Cfg::KSUB * (TC_PROD_WARPS / TC_GROUPS) * 32);   // every producer thread of the stage's K blocks
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < Cfg::B_STAGES; ++s) {
      mbar_init(bar_bfull + 8 * s, 1);
      mbar_init(bar_bempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, CG * TC_EPI_WARPS * 32);        // CG = 2: the epilogues of both CTAs
    }
    if constexpr (CG == 2)
      for (int s = 0; s < Cfg::A_STAGES; ++s) mbar_init(bar_pfull + 8 * s, 1);
    fence_mbar_init();
  }
  if (warp == TC_EPI_WARPS) {
    if constexpr (CG == 2) tmem_alloc2(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();             // barriers of both CTAs initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (warp < TC_EPI_WARPS) {
    // =========================== epilogue ===========================
    const int r = warp * 32 + lane;
    int it = 0;
    for (int tile = w_first; tile < w_total; tile += w_stride, ++it) {
      const int ptile = g.reverse ? w_total - 1 - tile : tile;
      const int m0 = (CG == 2 ? 2 * (ptile / p.n_tiles) + (int)cta_rank : ptile / p.n_tiles) * TC_BM;
      const int n0 = (ptile % p.n_tiles) * BN;
      const int as = it & 1;
      if (p.debug & 8192) mbar_wait(bar_tfull + 8 * as, (it >> 1) & 1);
      else mbar_wait_relaxed(bar_tfull + 8 * as, (it >> 1) & 1);
      tc_fence_after();
      const int m = m0 + r;
      const bool row_ok = m < g.M;
      const int64_t orow = row_ok ? (g.out_rows ? (int64_t)g.out_rows[m] : (int64_t)m) : 0;
      const float* radd = (row_ok && g.row_add) ? g.row_add + (int64_t)g.row_add_idx[m] * g.ld_row_add : nullptr;
      const __nv_bfloat16* res =
          (row_ok && g.resid) ? reinterpret_cast<const __nv_bfloat16*>(g.resid) + (int64_t)m * g.ld_resid : nullptr;
      constexpr int CH = BN >= 32 ? 32 : 16;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += CH) {
        uint32_t acc[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * BN + c0);
        if (CH == 32) { OF_TMEM_LD32(taddr, acc); } else { OF_TMEM_LD16(taddr, acc); }
        tmem_ld_wait();
        if (!row_ok || (p.debug & 4)) continue;
        const int nb = n0 + c0;
        float v[32];
#pragma unroll
        for (int j = 0; j < CH; ++j) v[j] = __uint_as_float(acc[j]);
        const bool full = (nb + CH <= g.N);
        if (g.bias) {
#pragma unroll
          for (int j = 0; j < CH; ++j) if (full || nb + j < g.N) v[j] += g.bias[nb + j];
        }
        if (radd) {
#pragma unroll
          for (int j = 0; j < CH; ++j) if (full || nb + j < g.N) v[j] += radd[nb + j];
        }
        if (res) {
          if (full && (g.ld_resid % 8 == 0)) {
#pragma unroll
            for (int q = 0; q < CH / 8; ++q) {
              float f[8];
              bf16x8_to_f32(ldg_nc_v4(res + nb + q * 8), f);
#pragma unroll
              for (int j = 0; j < 8; ++j) v[q * 8 + j] += f[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) if (nb + j < g.N) v[j] += __bfloat162float(res[nb + j]);
          }
        }
        if (g.out_f32) {
          float* o = reinterpret_cast<float*>(g.out) + orow * g.ldo + nb;
          if (full && (g.ldo % 4 == 0)) {
#pragma unroll
            for (int q = 0; q < CH / 4; ++q)
              *reinterpret_cast<float4*>(o + q * 4) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) if (nb + j < g.N) o[j] = v[j];
          }
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(g.out) + orow * g.ldo + nb;
          if (full && (g.ldo % 8 == 0)) {
#pragma unroll
            for (int q = 0; q < CH / 8; ++q) *reinterpret_cast<uint4*>(o + q * 8) = f32_to_bf16x8(v + q * 8);
          } else {
#pragma unroll
            for (int j = 0; j < CH; ++j) if (nb + j < g.N) o[j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      tc_fence_before();
      if constexpr (CG == 2) {
        // the accumulator pair is released on the LEADER's barrier (it issues the MMAs that overwrite both halves)
        if (cta_rank == 0) mbar_arrive(bar_tempty + 8 * as);
        else mbar_arrive_cluster(map_to_cta(bar_tempty + 8 * as, 0));
      } else {
        mbar_arrive(bar_tempty + 8 * as);
      }
    }
  } else if (warp == TC_EPI_WARPS) {
    // =========================== MMA issuer ===========================
    // the whole warp walks the pipeline (all lanes wait on the barriers); one elected lane issues
    if (CG == 2 && cta_rank != 0) {
      // pair, rank 1: no MMAs here.  Relay "my A stage and my half of B are full" to the leader's peer-ready ring;
      // the stages are released by the leader's multicast commits.
      int stage = 0, bstage = 0;
      uint32_t phase = 0, bphase = 0;
      for (int tile = w_first; tile < w_total; tile += w_stride) {
        for (int kb = 0; kb < p.num_kb; kb += Cfg::KSUB) {
          mbar_wait(bar_bfull + 8 * bstage, bphase);
          mbar_wait(bar_full + 8 * stage, phase);
          if (lane == 0) mbar_arrive_cluster(map_to_cta(bar_pfull + 8 * stage, 0));
          __syncwarp();
          if (++stage == Cfg::A_STAGES) { stage = 0; phase ^= 1; }
          if (++bstage == Cfg::B_STAGES) { bstage = 0; bphase ^= 1; }
        }
      }
    } else {
      constexpr uint32_t idesc = CG == 2 ? make_idesc_pair(BN) : make_idesc(BN);
      int stage = 0, bstage = 0;
      uint32_t phase = 0, bphase = 0;
      int it = 0;
      for (int tile = w_first; tile < w_total; tile += w_stride, ++it) {
        const int as = it & 1;
        mbar_wait(bar_tempty + 8 * as, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        for (int kb = 0; kb < p.num_kb; kb += Cfg::KSUB) {
          if (!(p.debug & 64)) mbar_wait(bar_bfull + 8 * bstage, bphase);
          mbar_wait(bar_full + 8 * stage, phase);
          if constexpr (CG == 2) mbar_wait(bar_pfull + 8 * stage, phase);   // ... and the peer's stages
          if (p.debug & 256) tc_fence_after();     // (experiment) not needed: the operands arrive by cp.async / TMA, not tcgen05
          const uint32_t a_addr = stage_base + stage * Cfg::A_BYTES;
          const uint32_t b_addr = b_ring + bstage * Cfg::B_BYTES;
          if (elect_one()) {
            // descriptor low words: start address >> 4 (+2 per 32-byte K step), LBO = 1; the high word is constant
            const uint32_t a_lo = ((a_addr & 0x3FFFFu) >> 4) | (1u << 16);
            const uint32_t b_lo = ((b_addr & 0x3FFFFu) >> 4) | (1u << 16);
            if (!(p.debug & 8)) {
#pragma unroll
              for (int j = 0; j < Cfg::KSUB; ++j) {
                if (j > 0 && kb + j >= p.num_kb) break;              // odd K-block count: the last stage is half full
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {
                  if constexpr (CG == 2)
                    umma_bf16_lo2(d_tmem, a_lo + j * (Cfg::A_SUB_BYTES >> 4) + 2 * k,
                                  b_lo + j * (Cfg::B_SUB_BYTES >> 4) + 2 * k, idesc, (kb > 0 || j > 0 || k > 0) ? 1u : 0u);
                  else
                    umma_bf16_lo(d_tmem, a_lo + j * (Cfg::A_SUB_BYTES >> 4) + 2 * k, b_lo + j * (Cfg::B_SUB_BYTES >> 4) + 2 * k,
                                 idesc, (kb > 0 || j > 0 || k > 0) ? 1u : 0u);
                }
              }
            }
            if constexpr (CG == 2) {                       // multicast: the same barrier offsets in both CTAs
              umma_commit2(bar_empty + 8 * stage);
              umma_commit2(bar_bempty + 8 * bstage);
              if (kb + Cfg::KSUB >= p.num_kb) umma_commit2(bar_tfull + 8 * as);
            } else {
              umma_commit(bar_empty + 8 * stage);            // frees the A stage when these MMAs retire
              if (!(p.debug & 64)) umma_commit(bar_bempty + 8 * bstage);          // ... and the B stage
              if (kb + Cfg::KSUB >= p.num_kb) umma_commit(bar_tfull + 8 * as);   // accumulator complete -> epilogue
            }
          }
          __syncwarp();
          if (++stage == Cfg::A_STAGES) { stage = 0; phase ^= 1; }
          if (++bstage == Cfg::B_STAGES) { bstage = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp == TC_EPI_WARPS + 1) {
    // =========================== weight loader ===========================
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint8_t* wp = reinterpret_cast<const uint8_t*>(g.w);
      for (int tile = w_first; tile < w_total; tile += w_stride) {
        // CG = 2: this CTA stages rows [rank * BN/2, +BN/2) of the [BN x 64] tile (a contiguous half of the packed image)
        const int n0 = ((g.reverse ? w_total - 1 - tile : tile) % p.n_tiles) * BN + (CG == 2 ? (int)cta_rank * (BN / 2) : 0);
        for (int kb = 0; kb < p.num_kb && !(p.debug & 64); kb += Cfg::KSUB) {
          mbar_wait(bar_bempty + 8 * stage, phase ^ 1);
          const uint32_t b_addr = b_ring + stage * Cfg::B_BYTES;
          if (elect_one()) {
            if (p.debug & 2) { mbar_arrive(bar_bfull + 8 * stage); }
            else {
              const int nk = min(Cfg::KSUB, p.num_kb - kb);
              mbar_arrive_expect_tx(bar_bfull + 8 * stage, (uint32_t)nk * Cfg::B_SUB_BYTES);
              for (int j = 0; j < nk; ++j)
                bulk_g2s(b_addr + j * Cfg::B_SUB_BYTES, wp + ((int64_t)(kb + j) * p.npad + n0) * 128, Cfg::B_SUB_BYTES,
                         bar_bfull + 8 * stage);
            }
          }
          __syncwarp();
          if (++stage == Cfg::B_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =========================== gather producers ===========================
    // 4 independent groups of 2 warps; group g produces the K blocks whose running index is = g mod 4,
    // so 4 K blocks (64 KB of gathers) are in flight per SM and the memory latency of one block is
    // hidden behind the other three.  Each thread owns one 16-byte chunk column (q) of 16 rows.
    const int pt = threadIdx.x - (TC_EPI_WARPS + 2) * 32;           // 0..255
    const int grp = pt >> 6;                                        // producer group 0..3
    const int gt = pt & 63;
    const int q = gt & 7;                                           // 16-byte chunk of the 128-byte row
    const int rbase = gt >> 3;                                      // 0..7, rows rbase + 8*i
    const __nv_bfloat16* a0 = reinterpret_cast<const __nv_bfloat16*>(g.a0);
    const __nv_bfloat16* a1 = reinterpret_cast<const __nv_bfloat16*>(g.a1);
    const int32_t* __restrict__ tab = g.tap_tab;
    // K blocks of this CTA in consumption order: kbg = tile_iter * num_kb + kb; this group owns kbg = grp (mod 4).
    // The 16 table entries of the NEXT owned block are fetched before waiting for the current stage to be
    // released, which takes the table latency off the stage turnaround.
    const int my_tiles = (w_total - w_first + w_stride - 1) / w_stride;
    // K-block slots of a tile, padded to whole stages (a pad slot carries no data: its owner only arrives)
    const uint32_t slots = (uint32_t)((p.num_kb + Cfg::KSUB - 1) / Cfg::KSUB * Cfg::KSUB);
    const uint32_t kb_total = (uint32_t)my_tiles * slots;
    const int feat_kb = p.cblocks * taps;
    // position of a K-block slot inside this CTA's work: (tile iteration, K block, channel block, tap), advanced
    // incrementally -- no integer divisions in the producer loop (its instruction stream competes with the MMA warp)
    struct Pos { int ti, kb, cb, tap; };
    auto norm = [&](Pos& s) {
      while (s.kb >= (int)slots) { s.kb -= (int)slots; ++s.ti; s.cb = 0; s.tap = s.kb; }
      while (s.tap >= taps) { s.tap -= taps; ++s.cb; }
    };
    auto tile_m0 = [&](int ti) {
      const int tile = w_first + ti * w_stride;
      const int pt = g.reverse ? w_total - 1 - tile : tile;
      const int mt = p.n_tiles == 1 ? pt : pt / p.n_tiles;
      return (CG == 2 ? 2 * mt + (int)cta_rank : mt) * TC_BM;
    };
    auto fetch_taps = [&](const Pos& s, int32_t* t) {
      const int kb = s.kb;
      const int m0 = tile_m0(s.ti);
      if (kb >= feat_kb || (p.debug & 32)) return;
      const int tap = s.tap;
      if (tab != nullptr) {
        const uint32_t base = (uint32_t)(m0 + rbase) * (uint32_t)taps + (uint32_t)tap;     // < 2^31 (checked on host)
#pragma unroll
        for (int i = 0; i < TC_BM / 8; ++i) {
          const int m = m0 + rbase + 8 * i;
          t[i] = -1;
          if (m < g.M) t[i] = __ldg(tab + (base + (uint32_t)(8 * i) * (uint32_t)taps));
        }
      } else if (g.in_rows != nullptr) {
#pragma unroll
        for (int i = 0; i < TC_BM / 8; ++i) {
          const int m = m0 + rbase + 8 * i;
          t[i] = -1;
          if (m < g.M) t[i] = __ldg(g.in_rows + m);
        }
      } else {
#pragma unroll
        for (int i = 0; i < TC_BM / 8; ++i) {
          const int m = m0 + rbase + 8 * i;
          t[i] = m < g.M ? m : -1;
        }
      }
    };
    int32_t tnext[TC_BM / 8];
    Pos cur{0, grp, 0, grp};
    norm(cur);
    if ((uint32_t)grp < kb_total) fetch_taps(cur, tnext);
    for (uint32_t kbg = (uint32_t)grp; kbg < kb_total; kbg += TC_GROUPS) {
      {
        const int kb = cur.kb;
        const int cur_cb = cur.cb;
        const int m0 = tile_m0(cur.ti);
        int32_t t[TC_BM / 8];
#pragma unroll
        for (int i = 0; i < TC_BM / 8; ++i) t[i] = tnext[i];
        cur.kb += TC_GROUPS; cur.tap += TC_GROUPS;
        norm(cur);                                           // now the position of kbg + TC_GROUPS
        if (kbg + TC_GROUPS < kb_total) fetch_taps(cur, tnext);
        const uint32_t sg = kbg / Cfg::KSUB;                       // stage counter; this K block is its sub-tile kbg % KSUB
        const uint32_t stage = sg % Cfg::A_STAGES;
        const uint32_t phase = (sg / Cfg::A_STAGES) & 1u;
        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
        const uint32_t a_addr = stage_base + stage * Cfg::A_BYTES + (kbg % Cfg::KSUB) * Cfg::A_SUB_BYTES;
        if ((p.debug & 1) || kb >= p.num_kb) {
        } else if (kb < p.cblocks * taps) {
          const int cb = cur_cb;
          const int ch = cb * TC_BK;
          const __nv_bfloat16* src;
          int64_t ld;
          if (ch < g.c0) { src = a0 + ch; ld = g.lda0; } else { src = a1 + (ch - g.c0); ld = g.lda1; }
          src += q * 8;
          if (p.debug & 16) {
#pragma unroll
            for (int i = 0; i < TC_BM / 8; ++i) t[i] = 0;          // timing experiment: pure L1 hits
          }
          const __nv_bfloat16* msrc = reinterpret_cast<const __nv_bfloat16*>(g.a_multi) + ch + q * 8;
          // Two independent data paths share the gather.  Rows 8i+0..3 (i = 0..15) of the tile: TMA gather4 --
          // one instruction moves four source rows x 128 B and swizzles in hardware; the four row indices are
          // held by lanes q, q+8, q+16, q+24 of the group's first warp.  Rows 8i+4..7: 16-byte cp.async
          // (LDGSTS) through L1.  Either path alone saturates at ~16 KB/us per SM (request queues), together
          // they overlap.
          const bool tma_rows = p.use_tma && rbase < 4;           // this thread's rows belong to the TMA half
          if (p.use_tma && gt < 32) {                             // first warp of the group: all 32 lanes converge
            const CUtensorMap* tm = ch < g.c0 ? &tmap0 : &tmap1;
            const int col = ch < g.c0 ? ch : ch - g.c0;
            const int oob = ch < g.c0 ? p.rows0 : p.rows1;
#pragma unroll
            for (int i = 0; i < TC_BM / 8; ++i) {
              const int t1 = __shfl_sync(0xffffffffu, t[i], q + 8);
              const int t2 = __shfl_sync(0xffffffffu, t[i], q + 16);
              const int t3 = __shfl_sync(0xffffffffu, t[i], q + 24);
              if (rbase == 0 && (i & 7) == q) {                   // lane q issues groups i = q and q + 8
                const int t0 = t[i];
                const uint32_t dst = a_addr + (8 * i) * 128;
                if (t0 >= -1 && t1 >= -1 && t2 >= -1 && t3 >= -1) {
                  mbar_expect_tx(bar_full + 8 * stage, 512u);
                  tma_gather4(dst, tm, col, t0 < 0 ? oob : t0, t1 < 0 ? oob : t1, t2 < 0 ? oob : t2, t3 < 0 ? oob : t3,
                              bar_full + 8 * stage);
                } else {
                  // a multi-neighbour slot in the group: its pre-averaged row lives in another tensor -> LDGSTS
                  const int tt[4] = {t0, t1, t2, t3};
                  const __nv_bfloat16* s0 = src - q * 8;
                  const __nv_bfloat16* m0p = msrc - q * 8;
#pragma unroll
                  for (int r4 = 0; r4 < 4; ++r4) {
                    const int rr = 8 * i + r4;
                    const int tv = tt[r4];
                    const __nv_bfloat16* base = tv >= 0 ? s0 + (int64_t)tv * ld
                                                        : (tv == -1 ? s0 : m0p + (int64_t)(-(tv + 2)) * g.ld_multi);
#pragma unroll
                    for (int c8 = 0; c8 < 8; ++c8)
                      cp_async_16(a_addr + rr * 128 + ((c8 ^ (rr & 7)) << 4), base + c8 * 8, tv == -1 ? 0u : 16u);
                  }
                }
              }
            }
          }
          if (!tma_rows) {
            // 16 asynchronous 16-byte global->shared copies back to back (no registers, no waiting):
            // one neighbour -> its row; none -> zero fill; several -> the pre-averaged row of a_multi.
            // Branch-free address: one select of (base, stride) + one 32x32+64 multiply-add per copy -- the producers'
            // instruction stream is what the MMA warp competes with for issue slots.
            const uint32_t dst0 = a_addr + rbase * 128 + ((q ^ (rbase & 7)) << 4);    // (rbase + 8i) & 7 == rbase & 7
            const uint64_t sbase = reinterpret_cast<uint64_t>(src), mbase = reinterpret_cast<uint64_t>(msrc);
            const uint32_t ldb = (uint32_t)ld * 2u, ldmb = (uint32_t)g.ld_multi * 2u;  // row strides in bytes
            // Multi-neighbour slots only occur on rows of coarse leaves (the first rows of the graph): two thirds of
            // the tiles have none, and then every copy is max / multiply-add / compare / LDGSTS.
            int32_t lo = t[0];
#pragma unroll
            for (int i = 1; i < TC_BM / 8; ++i) lo = min(lo, t[i]);
            if (!__any_sync(0xffffffffu, lo < -1)) {
#pragma unroll
              for (int i = 0; i < TC_BM / 8; ++i) {
                const int32_t tv = t[i];
                const uint64_t addr = sbase + (uint64_t)(uint32_t)max(tv, 0) * (uint64_t)ldb;
                cp_async_16(dst0 + i * 1024, reinterpret_cast<const void*>(addr), tv == -1 ? 0u : 16u);
              }
            } else {
#pragma unroll
              for (int i = 0; i < TC_BM / 8; ++i) {
                const int32_t tv = t[i];
                const bool multi = tv < -1;
                const uint32_t idx = multi ? (uint32_t)(-2 - tv) : (uint32_t)(tv < 0 ? 0 : tv);
                const uint64_t addr = (multi ? mbase : sbase) + (uint64_t)idx * (uint64_t)(multi ? ldmb : ldb);
                cp_async_16(dst0 + i * 1024, reinterpret_cast<const void*>(addr), tv == -1 ? 0u : 16u);
              }
            }
          }
        } else if (g.nt_block != nullptr) {
          // node-type block, precomputed per graph (of_graph_type_block): a plain coalesced copy of rows m0..m0+127
          const __nv_bfloat16* nb = reinterpret_cast<const __nv_bfloat16*>(g.nt_block) + q * 8;
#pragma unroll
          for (int i = 0; i < TC_BM / 8; ++i) {
            const int rr = rbase + 8 * i;
            const int m = m0 + rr;
            cp_async_16(a_addr + rr * 128 + ((q ^ (rr & 7)) << 4), m < g.M ? (const void*)(nb + (int64_t)m * 64) : (const void*)nb,
                        m < g.M ? 16u : 0u);
          }
        } else {
          // node-type block: column tap*ntype + type holds (#neighbours of that type)/(#neighbours)
          // = mean of the one-hot columns the reference concatenates (modules.py:199-202).
#pragma unroll 1
          for (int rr = gt; rr < TC_BM; rr += 64) {
            const uint32_t rowaddr = a_addr + rr * 128;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int c = 0; c < 8; ++c) sts_v4(rowaddr + (c << 4), z);
            const int m = m0 + rr;
            if (m >= g.M) continue;
            for (int tap = 0; tap < taps; ++tap) {
              const int32_t tv = tab ? __ldg(tab + (int64_t)m * taps + tap) : m;
              if (tv == -1) continue;
              unsigned long long packed;
              int n = 1;
              if (tv >= 0) {
                packed = 1ull << (8 * g.node_type[tv]);
              } else {
                packed = g.multi_types[-(tv + 2)];           // per-type neighbour counts of the slot
                n = 0;
                for (int ty = 0; ty < 8; ++ty) n += (int)((packed >> (8 * ty)) & 255ull);
              }
              for (int ty = 0; ty < g.ntype && ty < 8; ++ty) {
                const int c = (int)((packed >> (8 * ty)) & 255ull);
                if (c == 0) continue;
                const int col = tap * g.ntype + ty;
                const __nv_bfloat16 hv = __float2bfloat16_rn((float)c / (float)n);
                sts_u16(rowaddr + ((((col >> 3) ^ (rr & 7)) << 4) | ((col & 7) << 1)), __bfloat16_as_ushort(hv));
              }
            }
          }
          fence_proxy_async_smem();               // generic-proxy stores -> visible to the tensor core
        }
        // CUTLASS sm100 cp.async+UMMA protocol: one arrive that fires when this thread's cp.asyncs have
        // landed (self-incrementing, not counted) + one ordinary release-arrive (counted)
        // one counted arrival per thread: for gathered blocks it fires when this thread's cp.asyncs have landed
        // (cp.async.mbarrier.arrive.noinc); for the node-type block (generic stores + proxy fence) a plain arrive
        if ((kb < p.cblocks * taps || g.nt_block != nullptr) && kb < p.num_kb && !(p.debug & 1)) cp_async_mbar_arrive_noinc(bar_full + 8 * stage);
        else mbar_arrive(bar_full + 8 * stage);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();             // the peer may still be reading its half of the accumulators
  if (warp == TC_EPI_WARPS) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing: canonical fp32 [taps*(c+ntype), N] -> bf16 image [num_kb][npad rows][64], each
// 8-row group 128B-swizzled exactly as the MMA expects it, so a [BN x 64] tile is one contiguous
// cp.async.bulk.  K-block order = (channel block outer, tap inner), then the node-type block.
// ------------------------------------------------------------------------------------------------
__global__ void pack_weight_tc_kernel(const float* __restrict__ w, int taps, int c, int ntype, int N, int npad,
                                      int num_kb, __nv_bfloat16* __restrict__ out) {
  const int64_t total = (int64_t)num_kb * npad * 64;
  const int cblocks = c / 64;
  const int cp = c + ntype;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i & 63);                       // logical k within the block
    const int64_t rn = i >> 6;
    const int n = (int)(rn % npad);
    const int kb = (int)(rn / npad);
    float v = 0.0f;
    if (n < N) {
      if (kb < cblocks * taps) {
        const int cb = kb / taps, tap = kb - cb * taps;
        v = w[((int64_t)tap * cp + cb * 64 + j) * N + n];
      } else if (j < taps * ntype) {
        const int tap = j / ntype, ty = j - tap * ntype;
        v = w[((int64_t)tap * cp + c + ty) * N + n];
      }
    }
    const int64_t dst = ((int64_t)kb * npad + n) * 64 + ((((j >> 3) ^ (n & 7)) << 3) | (j & 7));
    out[dst] = __float2bfloat16_rn(v);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// 2-D bf16 tensor [rows, c] with row stride ld (elements); box = 64 columns x 1 row (the gather4 unit), 128B swizzle
static bool make_row_tmap(CUtensorMap* m, const void* base, int64_t rows, int c, int64_t ld) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn || base == nullptr || rows <= 0) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)c, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, 1};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// EXPERIMENTAL CTA-pair launch (OCTFUSION_TC_CTA2=1): cluster of 2, even grid, BN = 256 only
template <int BN>
static int launch_tc_pair(const TcParams& p, const CUtensorMap& t0, const CUtensorMap& t1, cudaStream_t st) {
  using Cfg = TcCfg<BN, 2>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gather_gemm_tc_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("of_gather_gemm_tc (pair): cudaFuncSetAttribute(%d B): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return OF_E_CUDA;
    }
    configured = true;
  }
  const int pair_tiles = ((p.m_tiles + 1) / 2) * p.n_tiles;
  int grid = 2 * pair_tiles < num_sms() ? 2 * pair_tiles : (num_sms() & ~1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gather_gemm_tc_pair_kernel<BN>, p, t0, t1);
  if (e != cudaSuccess) {
    set_error("of_gather_gemm_tc (pair): launch: %s", cudaGetErrorString(e));
    return OF_E_CUDA;
  }
  OF_LAUNCH_CHECK("of_gather_gemm_tc(pair)");
  return OF_OK;
}

template <int BN>
static int launch_tc(const TcParams& p, const CUtensorMap& t0, const CUtensorMap& t1, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gather_gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("of_gather_gemm_tc: cudaFuncSetAttribute(%d B): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return OF_E_CUDA;
    }
    configured = true;
  }
  const int total = p.m_tiles * p.n_tiles;
  int grid = total < num_sms() ? total : num_sms();
  {
    static int lim = -1;                                   // OCTFUSION_TC_GRID: cap the CTA count (experiments)
    if (lim < 0) { const char* e = getenv("OCTFUSION_TC_GRID"); lim = e ? atoi(e) : 0; }
    if (lim > 0 && grid > lim) grid = lim;
  }
  gather_gemm_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(p, t0, t1);
  OF_LAUNCH_CHECK("of_gather_gemm_tc");
  return OF_OK;
}

int check_gemm_args(const of_gemm_args* a, const char* who);

}  // namespace of

using namespace of;

extern "C" int64_t of_pack_weight_tc_bytes(int32_t taps, int32_t c, int32_t ntype, int32_t N) {
  if (taps <= 0 || c <= 0 || c % 64 != 0 || N <= 0 || ntype < 0 || taps * ntype > 64) return -1;
  const int64_t num_kb = (int64_t)taps * (c / 64) + (ntype > 0 ? 1 : 0);
  const int64_t npad = (N + 15) / 16 * 16;
  return num_kb * npad * 64 * 2;
}

extern "C" int of_pack_weight_tc(const float* w_canonical, int32_t taps, int32_t c, int32_t ntype, int32_t N,
                                 void* out, void* stream) {
  OF_REQUIRE(w_canonical && out, "of_pack_weight_tc: null pointer");
  OF_REQUIRE(of_pack_weight_tc_bytes(taps, c, ntype, N) > 0, "of_pack_weight_tc: unsupported shape taps=%d c=%d nt=%d N=%d",
             taps, c, ntype, N);
  const int num_kb = taps * (c / 64) + (ntype > 0 ? 1 : 0);
  const int npad = (N + 15) / 16 * 16;
  const int64_t total = (int64_t)num_kb * npad * 64;
  int64_t want = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 32;
  const int grid = (int)(want < cap ? want : cap);
  pack_weight_tc_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      w_canonical, taps, c, ntype, N, npad, num_kb, reinterpret_cast<__nv_bfloat16*>(out));
  OF_LAUNCH_CHECK("of_pack_weight_tc");
  return OF_OK;
}

extern "C" int of_gather_gemm_tc(const of_gemm_args* args, void* stream) {
  int rc = check_gemm_args(args, "of_gather_gemm_tc");
  if (rc) return rc;
  const of_gemm_args& a = *args;
  if (a.dtype != OF_BF16 || a.c0 % 64 != 0 || a.c1 % 64 != 0 || a.taps > TC_MAX_TAPS || a.taps * a.ntype > 64 ||
      a.ntype > 8 || a.a_silu) {
    set_error("of_gather_gemm_tc: unsupported (dtype=%d c0=%d c1=%d taps=%d ntype=%d a_silu=%d)", a.dtype, a.c0, a.c1,
              a.taps, a.ntype, a.a_silu);
    return OF_E_UNSUPPORTED;
  }
  OF_REQUIRE(a.lda0 % 8 == 0 && (a.c1 == 0 || a.lda1 % 8 == 0), "of_gather_gemm_tc: lda must be a multiple of 8");
  OF_REQUIRE(reinterpret_cast<uintptr_t>(a.a0) % 16 == 0 && reinterpret_cast<uintptr_t>(a.a1) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(a.w) % 16 == 0,
             "of_gather_gemm_tc: a0/a1/w must be 16-byte aligned");
  if (a.M == 0) return OF_OK;
  TcParams p;
  p.g = a;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("OCTFUSION_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
  }
  p.cblocks = (a.c0 + a.c1) / 64;
  p.num_kb = p.cblocks * a.taps + (a.ntype > 0 ? 1 : 0);
  p.npad = (a.N + 15) / 16 * 16;
  p.m_tiles = (a.M + TC_BM - 1) / TC_BM;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // TMA gather4 path: needs the row counts of the sources (out-of-bounds row = zero fill) and a tensor map each
  alignas(64) CUtensorMap t0, t1;
  memset(&t0, 0, sizeof(t0)); memset(&t1, 0, sizeof(t1));
  p.use_tma = 0; p.rows0 = a.rows_a0; p.rows1 = a.rows_a1;
  {
    // Measured on B200 (profiles/tc_gather_experiments_r01.md): routing half of the rows through TMA gather4
    // is ~1.6x SLOWER than 16-byte cp.async for this access pattern, so it is opt-in (OCTFUSION_TC_TMA=1).
    static int want_tma = -1;
    if (want_tma < 0) { const char* e = getenv("OCTFUSION_TC_TMA"); want_tma = e ? atoi(e) : 0; }
    if (want_tma && a.rows_a0 > 0 && (a.c1 == 0 || a.rows_a1 > 0) && make_row_tmap(&t0, a.a0, a.rows_a0, a.c0, a.lda0) &&
        (a.c1 == 0 || make_row_tmap(&t1, a.a1, a.rows_a1, a.c1, a.lda1)))
      p.use_tma = 1;
  }
  // widest tile that divides the padded N: fewer re-gathers of A per output column
  if (p.npad % 256 == 0) {
    p.n_tiles = p.npad / 256;
    static int want_pair = -1;                              // EXPERIMENTAL cta_group::2 variant, not yet run on hardware
    if (want_pair < 0) { const char* e = getenv("OCTFUSION_TC_CTA2"); want_pair = e ? atoi(e) : 0; }
    if (want_pair && !p.use_tma) return launch_tc_pair<256>(p, t0, t1, st);
    return launch_tc<256>(p, t0, t1, st);
  }
  if (p.npad % 128 == 0) { p.n_tiles = p.npad / 128; return launch_tc<128>(p, t0, t1, st); }
  if (p.npad % 64 == 0)  { p.n_tiles = p.npad / 64;  return launch_tc<64>(p, t0, t1, st); }
  if (p.npad % 32 == 0)  { p.n_tiles = p.npad / 32;  return launch_tc<32>(p, t0, t1, st); }
  p.n_tiles = p.npad / 16;
  return launch_tc<16>(p, t0, t1, st);
}
