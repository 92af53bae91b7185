// Thin inline-PTX wrappers for the sm_100a features the NPP kernels use:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and
// the UMMA shared-memory / instruction descriptors.
//
// Bit layouts follow the PTX ISA "tcgen05" chapter (shared memory descriptor,
// instruction descriptor for .kind::f16).  Nothing here is NPP specific.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_fp16.h>

namespace npp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// wait that synchronises with release.cluster arrives of the peer CTA (generic-proxy data forwarded through smem)
__device__ __forceinline__ void mbar_wait_cluster_impl(uint64_t* bar, uint32_t parity, uint32_t line);
#ifdef NPP_HANG_DEBUG
#define mbar_wait_cluster(bar, parity) mbar_wait_cluster_impl(bar, parity, __LINE__)
#else
#define mbar_wait_cluster(bar, parity) mbar_wait_cluster_impl(bar, parity, 0)
#endif
__device__ __forceinline__ void mbar_wait_cluster_plain(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (ok == 0);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
#ifdef NPP_HANG_DEBUG
// Debug build only (-DNPP_HANG_DEBUG): every mbarrier wait writes (source line, barrier, parity, clock) into a per-plan
// device buffer before it starts spinning and marks the entry "passed" afterwards; the spin loop itself is the
// production one (instrumenting the loop hid the dead-lock this was written for).  The kernels publish the buffer
// pointer in the last 8 bytes of their dynamic shared memory (never reached by the carve-up); the host reads the buffer
// on a private stream while the kernels hang (npp_debug_state_dump).  One slot per (block, warp), written by lane 0.
__device__ __forceinline__ unsigned long long** npp_state_slot() {
  extern __shared__ uint8_t npp_dyn_smem[];
  uint32_t sz;
  asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(sz));
  return reinterpret_cast<unsigned long long**>(npp_dyn_smem + sz - 8);
}
__device__ __forceinline__ void npp_state_record(uint32_t line, uint32_t bar, uint32_t parity, uint32_t passed) {
  if ((threadIdx.x & 31) != 0) return;
  unsigned long long* b = *npp_state_slot();
  if (b == nullptr) return;
  b[(blockIdx.x & 255) * 16 + (threadIdx.x >> 5)] =
      ((unsigned long long)(passed & 1) << 63) | ((unsigned long long)(line & 0x7FFF) << 48) |
      ((unsigned long long)(bar & 0xFFFF) << 32) | ((unsigned long long)(parity & 1) << 31) |
      ((unsigned long long)(clock64() >> 10) & 0x7FFFFFFFull);
}
__device__ __forceinline__ void mbar_wait_impl(uint64_t* bar, uint32_t parity, uint32_t line) {
  npp_state_record(line, smem_u32(bar), parity, 0);
  while (!mbar_try_wait(bar, parity)) {
  }
  npp_state_record(line, smem_u32(bar), parity, 1);
}
#define mbar_wait(bar, parity) mbar_wait_impl(bar, parity, __LINE__)
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
#ifdef NPP_SPIN_SLEEP_NS
    __nanosleep(NPP_SPIN_SLEEP_NS);   // experiment: energy of the spinning warps under the power cap
#endif
  }
}
#endif
__device__ __forceinline__ void mbar_wait_cluster_impl(uint64_t* bar, uint32_t parity, uint32_t line) {
#ifdef NPP_HANG_DEBUG
  npp_state_record(line, smem_u32(bar), parity, 0);
  mbar_wait_cluster_plain(bar, parity);
  npp_state_record(line, smem_u32(bar), parity, 1);
#else
  (void)line;
  mbar_wait_cluster_plain(bar, parity);
#endif
}

// ---------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load, completion signalled on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// Same load, delivered to the same smem offset (and signalling the mbarrier at the same offset) in every CTA of the
// cluster whose bit is set in cta_mask.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
        "h"(cta_mask)
      : "memory");
}
// cta_group::2 load: data lands in the executing CTA's smem, the transaction bytes are credited to the mbarrier at
// the same offset in CTA 0 of the pair (the leader that issues the 2-SM UMMAs).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
#ifdef NPP_PAIR_MBAR_MAPA
  uint32_t lead_bar;   // experiment: the leader's barrier through mapa instead of clearing the rank bit of the own address
  asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(lead_bar) : "r"(smem_u32(bar)));
#else
  const uint32_t lead_bar = smem_u32(bar) & 0xFEFFFFFFu;
#endif
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(lead_bar), "r"(c0), "r"(c1)
      : "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `target_rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t target_rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n"
      :
      : "r"(smem_u32(bar)), "r"(target_rank)
      : "memory");
}
// Same without the release fence (which compiles to MEMBAR.ALL.GPU and costs ~2000 clocks per arrive).  For
// signals that do not publish generic-proxy global memory: "my tcgen05.ld of the accumulator are done" (ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync) and "my part of an smem operand is written" (ordered by the
// fence.proxy.async each lane executed before the __syncwarp that precedes this arrive).
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint64_t* bar, uint32_t target_rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}\n"
      :
      : "r"(smem_u32(bar)), "r"(target_rank)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-D tiled store smem -> global (bulk async group), OOB parts of the box are clipped.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING their smem source
// all but the most recent committed bulk store of this thread have finished READING their smem source
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the most recently committed bulk group of this thread have completed (writes performed)
__device__ __forceinline__ void bulk_wait2() { asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait1() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes -> visible to the async proxy (UMMA / TMA reads of smem)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// orders async-proxy accesses (TMA loads/stores) of ANY state space against generic-proxy accesses
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ------------------------------------------------------- programmatic dependent launch
// griddepcontrol.wait: blocks until every kernel this launch depends on has completed and its writes are visible
// (no-op when the kernel was launched without the programmatic-stream-serialization attribute).
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// the dependent kernel of this launch may start its prologue (it still waits for this grid in grid_dependency_wait)
__device__ __forceinline__ void grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// named barrier for a subset of the CTA's warps
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ------------------------------------------------------------------ tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
// 2-SM variants: issued by the same logical warp in both CTAs of the pair, with the same smem result offset
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], f16 inputs / f32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 2-SM UMMA (M = 256 across the CTA pair), issued by one thread of the leader CTA
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// same, arriving on the barrier at this offset in every CTA of the cluster selected by cta_mask
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
        "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
        "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
        "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// --------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (64 bit), SWIZZLE_128B canonical layouts:
//   [ 0,14) start address >> 4      [16,30) leading byte offset >> 4
//   [32,46) stride byte offset >> 4 [46,48) version = 1 (sm_100)
//   [61,64) layout type: 2 = SWIZZLE_128B
// K-major  tile (rows x 64 halves, 128 B per row, 8-row atoms of 1024 B):
//     LBO unused (1), SBO = 1024 (next 8-row group).  A K=16 slice is +32 B.
// MN-major tile (64 halves contiguous along M/N, one 128-B row per K index,
//     8 K-rows = 1024 B atom): LBO = byte stride between 64-wide M/N atoms,
//     SBO = 1024 (next group of 8 K rows).  A K=16 slice is +2048 B.
__host__ __device__ constexpr uint64_t umma_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint64_t hi_bits) {
  return hi_bits | static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
}
// Instruction descriptor for kind::f16, f16 x f16 -> f32:
//   [4,6) D format (1 = f32)  [7,10) A format (0 = f16)  [10,13) B format (0 = f16)
//   [15] A major (0 = K, 1 = MN)  [16] B major  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

}  // namespace npp
