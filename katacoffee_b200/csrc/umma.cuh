// Thin inline-PTX layer for sm_100a: mbarrier, bulk async copy (TMA engine, UBLKCP), tcgen05
// (TMEM alloc, UMMA, commit, TMEM load) and the descriptor encodings used by the trunk kernel.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 shared memory descriptor" / "instruction
// descriptor" tables (same encodings as cute/arch/mma_sm100_desc.hpp in CUTLASS):
//   smem descriptor: [0,14) start>>4, [16,30) leading byte offset>>4, [32,46) stride byte offset>>4,
//                    [46,48) version = 1, [49,52) base offset, [61,64) swizzle (0 = none)
//   instr descriptor (kind::f16): [4,6) D format (1 = f32), [7,10) A format (1 = bf16), [10,13) B format,
//                    bit 15 A major (0 = K), bit 16 B major (0 = K), [17,23) N>>3, [24,29) M>>4
// Canonical K-major, no-swizzle operand layout: 8x(16 byte) core matrices; rows of one 8-element
// K-chunk are 16 B apart (so 8-row groups are SBO = 128 B apart when rows are stored densely) and the
// two K-chunks of one K=16 MMA are LBO apart.  Dense rows make "row + s" a plain +16*s on the start
// address, which is what turns a 3x3 tap into a descriptor offset.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace kc {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
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
    : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// Bounded wait: gives up (and raises *abortFlag) after ~2^32 cycles so that a protocol bug ends the
// kernel with an error code instead of hanging the GPU.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abortFlag, int code) {
  if(mbar_try_wait(bar, parity)) return true;
  long long t0 = clock64();
  for(uint32_t it = 1;; it++) {
    if(mbar_try_wait(bar, parity)) return true;
    if((it & 255u) == 0) {
      if(*abortFlag != 0) return false;
      if(clock64() - t0 > (1LL << 32)) { *abortFlag = code; return false; }
    }
  }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// global -> shared bulk copy executed by the TMA engine; completion is signalled on `bar` (tx bytes)
__device__ __forceinline__ void bulk_g2s(uint32_t dstSmem, const void* srcGlobal, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dstSmem), "l"(srcGlobal), "r"(bytes), "r"(bar) : "memory");
}

// ---- TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t dstSmem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dstSmem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lboBytes, uint32_t sboBytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lboBytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sboBytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version for sm_100
  return d;                 // base offset 0, swizzle none
}
// kind::f16 with fp32 accumulation; aFmt / bFmt: 0 = fp16, 1 = bf16 (the two operands choose independently, same MMA rate)
__host__ __device__ __forceinline__ uint32_t idesc_f16kind_f32(int M, int N, uint32_t aFmt, uint32_t bFmt) {
  return (1u << 4) | (aFmt << 7) | (bFmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ __forceinline__ uint32_t idesc_bf16_f32(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, one thread issues
__device__ __forceinline__ void umma_bf16(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
    ::"r"(dTmem), "l"(aDesc), "l"(bDesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Weight-stationary form: the B operand is latched in collector buffer b<BUF>; FILL loads it from shared
// memory, a following USE/LASTUSE MMA re-uses the latched copy, so two MMAs that share B (the two
// activation tiles of a CTA against one weight block) read B from shared memory once.
// MODE: 0 = fill, 1 = use, 2 = lastuse, 3 = discard.  Shapes: M in {32,64,128}, N in {64,128,256}.
template <int BUF, int MODE>
__device__ __forceinline__ void umma_bf16_ws(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t idesc, uint32_t accumulate) {
#define KC_WS_ASM(SUFFIX)                                                                                   \
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"                                          \
               "tcgen05.mma.ws.cta_group::1.kind::f16.collector::" SUFFIX " [%0], %1, %2, %3, p;\n\t}"      \
               ::"r"(dTmem), "l"(aDesc), "l"(bDesc), "r"(idesc), "r"(accumulate) : "memory")
  if constexpr(BUF == 0 && MODE == 0) KC_WS_ASM("b0::fill");
  else if constexpr(BUF == 0 && MODE == 1) KC_WS_ASM("b0::use");
  else if constexpr(BUF == 0 && MODE == 2) KC_WS_ASM("b0::lastuse");
  else if constexpr(BUF == 0 && MODE == 3) KC_WS_ASM("b0::discard");
  else if constexpr(BUF == 1 && MODE == 0) KC_WS_ASM("b1::fill");
  else if constexpr(BUF == 1 && MODE == 2) KC_WS_ASM("b1::lastuse");
  else if constexpr(BUF == 2 && MODE == 0) KC_WS_ASM("b2::fill");
  else if constexpr(BUF == 2 && MODE == 2) KC_WS_ASM("b2::lastuse");
  else if constexpr(BUF == 3 && MODE == 0) KC_WS_ASM("b3::fill");
  else if constexpr(BUF == 3 && MODE == 2) KC_WS_ASM("b3::lastuse");
  else static_assert(BUF < 0, "unsupported collector combination");
#undef KC_WS_ASM
}
// arrives on `bar` when all tcgen05 ops issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread i of the warp gets lane (base lane + i)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
    : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for(int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// Split form for software pipelining: issue the load of the next 16 columns, do the math of the current ones, then
// wait.  The wait names the destination registers as read-write operands so that no use of them can be scheduled
// above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
    : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t r[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
    : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
    :: "memory");
}

// ---- CTA pair (cluster of two CTAs on one TPC, tcgen05 cta_group::2) ----
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
// arrive on an mbarrier of any CTA of the cluster (address from mapa).  Default (CTA-scope release) semantics on
// purpose: what the arrival orders is this CTA's own shared-memory writes (already fenced to the async proxy) against
// its OWN tensor core reading them for the pair's next MMA; a cluster-scope release costs microseconds per arrival.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t clusterAddr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(clusterAddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {   // acquire at cluster scope
  uint32_t ok;
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
    "selp.u32 %0, 1, 0, p;\n\t}"
    : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity, volatile int* abortFlag, int code) {
  if(mbar_try_wait_cluster(bar, parity)) return true;
  long long t0 = clock64();
  for(uint32_t it = 1;; it++) {
    if(mbar_try_wait_cluster(bar, parity)) return true;
    if((it & 255u) == 0) {
      if(*abortFlag != 0) return false;
      if(clock64() - t0 > (1LL << 32)) { *abortFlag = code; return false; }
    }
  }
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dstSmem, uint32_t ncols) {  // whole warp, in each CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dstSmem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A (128 rows from each CTA's shared memory) * B^T (N/2 rows from each CTA's shared memory);
// issued by one thread of the leader CTA (rank 0) for the pair
__device__ __forceinline__ void umma_bf16_2cta(uint32_t dTmem, uint64_t aDesc, uint64_t bDesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
    "{\n\t.reg .pred p;\n\t"
    "setp.ne.b32 p, %4, 0;\n\t"
    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
    ::"r"(dTmem), "l"(aDesc), "l"(bDesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier at this shared-memory offset in every CTA of `mask` once all prior MMAs of the pair completed
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace ptx
}  // namespace kc
