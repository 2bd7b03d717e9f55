// sm_100a device helpers: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05/TMEM, bf16 packing.
// Hand-written inline PTX; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdod {

typedef __nv_bfloat16 bf16;

#define SDOD_DEVICE __device__ __forceinline__

SDOD_DEVICE uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
SDOD_DEVICE uint32_t lane_id() { return threadIdx.x & 31; }

SDOD_DEVICE bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ mbarrier
SDOD_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SDOD_DEVICE void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SDOD_DEVICE void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SDOD_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
SDOD_DEVICE void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
SDOD_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred P;\n mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n selp.u32 %0, 1, 0, P;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug traps (launch error) instead of hanging the GPU.
SDOD_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();   // a failed try_wait already suspends for a while: this is seconds, not a hot loop
    }
}

// ------------------------------------------------------------------ TMA
SDOD_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
SDOD_DEVICE void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
SDOD_DEVICE void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
SDOD_DEVICE void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// tensor store shared -> global (bulk async-group completion); out-of-bounds parts of the box are clipped
SDOD_DEVICE void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// TMA reduction store: global[box] += smem[box] (element type and swizzle from the tensor map; fp32 add is performed at L2)
SDOD_DEVICE void tma_reduce_add_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
SDOD_DEVICE void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
// 1-D bulk copy global -> shared (bytes multiple of 16, both 16-B aligned)
SDOD_DEVICE void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// 1-D bulk copy shared -> global
SDOD_DEVICE void bulk_store_1d(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(reinterpret_cast<uint64_t>(dst)), "r"(smem_u32(src)), "r"(bytes)
                 : "memory");
}
SDOD_DEVICE void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SDOD_DEVICE void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
SDOD_DEVICE void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the most recent bulk group of this thread have finished reading their shared-memory source
SDOD_DEVICE void bulk_wait_read_but_last() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05 / TMEM
template <int kCols>
SDOD_DEVICE void tmem_alloc(uint32_t* smem_dst) {   // whole warp, .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int kCols>
SDOD_DEVICE void tmem_dealloc(uint32_t taddr) {      // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}
// ------------------------------------------------------------------ programmatic dependent launch (PDL)
// A kernel launched with the programmatic-serialization attribute may start while its predecessor is still draining:
// everything before griddep_wait() (barrier init, TMEM allocation, descriptor prefetch) overlaps the predecessor's tail;
// griddep_wait() returns once the predecessor grid has completed and its writes are visible.  griddep_launch() lets the
// successor start its own preamble.  Both are no-ops for ordinary launches.
SDOD_DEVICE void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
SDOD_DEVICE void griddep_launch() {
    // Early trigger only for grids that fit the chip in one go (<= 2 CTAs per SM): there the successor's preamble hides behind our
    // tail.  For multi-wave grids an early successor takes SM slots from our own later waves — measured r1, batch-8 step: +2 %
    // time with unconditional triggers, while the batch-2 step (mostly single-wave kernels) gains 2.4 %.
    if (gridDim.x * gridDim.y * gridDim.z <= 296u) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ------------------------------------------------------------------ CTA pairs (cta_group::2): two SMs of one TPC share an MMA
SDOD_DEVICE uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
SDOD_DEVICE void cluster_sync_all() {     // every thread of every CTA in the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
SDOD_DEVICE void cluster_sync_unaligned() {     // same barrier, callable from diverged warps (each thread arrives on its own)
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// Split cluster barrier without memory ordering (UCGABAR only).  barrier.cluster.arrive.release compiles to MEMBAR.ALL.GPU + ERRBAR: a warp
// then waits for every global store it has in flight, which the short GroupNorm kernels cannot afford twice per launch.
SDOD_DEVICE void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
SDOD_DEVICE void cluster_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
// Two / three floats into CTA `rank`'s shared memory at the address `local_dst` has here, completion (8 / 12 bytes) counted on that CTA's
// mbarrier `local_bar`: the data and the transaction count travel together (STAS), so an all-gather of per-CTA statistics needs no fence,
// no release/acquire cluster barrier and no exit barrier — the receiver waits on its own mbarrier and reads its own shared memory.
SDOD_DEVICE void st_async_f32(const void* local_dst, const void* local_bar, uint32_t rank, float a) {
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_dst)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(ra), "r"(__float_as_uint(a)), "r"(rb) : "memory");
}
SDOD_DEVICE void st_async_f32x2(const void* local_dst, const void* local_bar, uint32_t rank, float a, float b) {
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_dst)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                 ::"r"(ra), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(rb) : "memory");
}
SDOD_DEVICE void st_async_f32x4(const void* local_dst, const void* local_bar, uint32_t rank, float a, float b, float c, float d) {
    uint32_t ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_dst)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(smem_u32(local_bar)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                 ::"r"(ra), "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)), "r"(rb) : "memory");
}
SDOD_DEVICE uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// fp32 load from the shared memory of CTA `rank` of this cluster, at the address `local_ptr` has in this CTA (distributed shared memory)
SDOD_DEVICE float ld_dsmem_f32(const float* local_ptr, uint32_t rank) {
    uint32_t ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) in CTA `rank` of the cluster
SDOD_DEVICE uint32_t mapa_shared(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
template <int kCols>
SDOD_DEVICE void tmem_alloc_pair(uint32_t* smem_dst) {   // one warp in EACH CTA of the pair, same smem offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
SDOD_DEVICE void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(kCols) : "memory");
}
// TMA loads whose completion bytes are credited to an mbarrier of either CTA of the pair (bar_cluster_addr from mapa_shared)
SDOD_DEVICE void tma_load_3d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
SDOD_DEVICE void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// D[tmem of both CTAs] (+)= A[256 x K: 128 rows from each CTA's smem] * B[N x K: N/2 rows from each CTA's smem]; leader CTA only
SDOD_DEVICE void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// MMA completion -> arrive on the mbarrier at this smem offset in BOTH CTAs of the pair
SDOD_DEVICE void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3)) : "memory");
}

SDOD_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SDOD_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// MMA completion -> mbarrier arrive (implies fence::before_thread_sync)
SDOD_DEVICE void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate
SDOD_DEVICE void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
SDOD_DEVICE void tc_mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
SDOD_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
SDOD_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t)
SDOD_DEVICE void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
SDOD_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
SDOD_DEVICE void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
        : "r"(taddr)
        : "memory");
}
SDOD_DEVICE void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
SDOD_DEVICE void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
          "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}

// K-major operand tile in smem, rows of 128 B (64 bf16), SWIZZLE_128B, 8-row groups 1024 B apart.
// (cute mma_sm100_desc.hpp SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
//  version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2.)
SDOD_DEVICE uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;             // LBO (unused for swizzled K-major; canonical value 1)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;     // SBO = 8 rows * 128 B
    d |= static_cast<uint64_t>(1) << 46;             // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major.
// c_format[4,6)=1 (F32), a_format[7,10)=1 (BF16), b_format[10,13)=1, n>>3 at [17,23), m>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// ------------------------------------------------------------------ numerics helpers
SDOD_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
SDOD_DEVICE float2 unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
// x * sigmoid(x) with MUFU.EX2 + MUFU.RCP (5 instructions).  The IEEE division of the first version (x / (1 + __expf(-x))) expands to MUFU.RCP + FCHK +
// a Newton step + a slow-path branch: ncu showed the in-network GroupNorm at batch 32 instruction-bound (issue slots 59 %, DRAM 14 %,
// profiles/r02_gn_group_b32_source_top.txt).  ~2 ulp of fp32; saturates correctly (x -> -inf gives -0, x -> +inf gives x).
SDOD_DEVICE float silu_f(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
SDOD_DEVICE float quick_gelu_f(float x) { return x / (1.0f + __expf(-1.702f * x)); }     // CLIP's QuickGELU: x * sigmoid(1.702 x)
// exact-erf GELU with erf from Abramowitz & Stegun 7.1.26 (|abs err| <= 1.5e-7): a handful of FMAs + one MUFU exp + one rcp
// instead of erff's long polynomial — the GEGLU epilogue evaluates it on 21 M elements per 64x64 transformer block.
SDOD_DEVICE float erf_fast(float x) {
    const float ax = fabsf(x);
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));   // 1 ulp; the IEEE form is a 6-instruction sequence
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));
    const float y = fmaf(-p * t, e, 1.0f);
    return copysignf(y, x);
}
SDOD_DEVICE float gelu_f(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752440f)); }

// Packed fp32 pairs (sm_100 FFMA2 / FADD2): two lanes of fp32 math per issue slot.
SDOD_DEVICE uint64_t pack_f32x2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
SDOD_DEVICE void unpack_f32x2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
SDOD_DEVICE uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
SDOD_DEVICE uint64_t add_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

SDOD_DEVICE uint64_t mul_f32x2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// Two exact-erf GELUs at once on the packed-fp32 pipe (same A&S 7.1.26 form as gelu_f; the polynomial carries negated
// coefficients so that no operand negation is needed): ~11 issue slots per element instead of ~19.
SDOD_DEVICE uint64_t gelu_f32x2(uint64_t x) {
    const uint64_t z = mul_f32x2(x, pack_f32x2(0.70710678118654752440f, 0.70710678118654752440f));
    const uint64_t az = z & 0x7fffffff7fffffffull;
    float d0, d1;
    unpack_f32x2(fma_f32x2(pack_f32x2(0.3275911f, 0.3275911f), az, pack_f32x2(1.0f, 1.0f)), d0, d1);
    float t0, t1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
    const uint64_t t = pack_f32x2(t0, t1);
    uint64_t q = fma_f32x2(pack_f32x2(-1.061405429f, -1.061405429f), t, pack_f32x2(1.453152027f, 1.453152027f));
    q = fma_f32x2(q, t, pack_f32x2(-1.421413741f, -1.421413741f));
    q = fma_f32x2(q, t, pack_f32x2(0.284496736f, 0.284496736f));
    q = fma_f32x2(q, t, pack_f32x2(-0.254829592f, -0.254829592f));
    q = mul_f32x2(q, t);                                                   // -(a1 t + ... + a5 t^5)
    float a0, a1;
    unpack_f32x2(mul_f32x2(mul_f32x2(az, pack_f32x2(1.2011224087864498f, 1.2011224087864498f)),
                           mul_f32x2(az, pack_f32x2(-1.2011224087864498f, -1.2011224087864498f))), a0, a1);   // -z^2 log2(e)
    float e0, e1;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
    const uint64_t y = fma_f32x2(q, pack_f32x2(e0, e1), pack_f32x2(1.0f, 1.0f));                             // erf(|z|)
    const uint64_t erf2 = y | (z & 0x8000000080000000ull);                                                   // y >= 0: copy z's sign in
    const uint64_t hx = mul_f32x2(x, pack_f32x2(0.5f, 0.5f));
    return fma_f32x2(hx, erf2, hx);
}

// GELU in its tanh form, 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))): five packed-fp32 operations and one MUFU.TANH per element against
// ~14 + two MUFU ops for the erf form above.  It differs from the exact GELU by < 5e-4 absolute (below one bf16 ulp of the product it feeds);
// on the batch-2 UNet step the switch moves eps by 3.9e-5 relative L2 against the 7.1e-3 that bf16 operands cost (tools/eps_error_budget.py).
// Opt-in only (SDOD_GEGLU_TANH=1): it buys 5 % on the isolated GEGLU projection and nothing measurable on the UNet pass.
SDOD_DEVICE uint64_t gelu_tanh_f32x2(uint64_t x) {
    const uint64_t x2 = mul_f32x2(x, x);
    const uint64_t p = fma_f32x2(x2, pack_f32x2(0.0356774081f, 0.0356774081f), pack_f32x2(0.7978845608f, 0.7978845608f));   // sqrt(2/pi) (1 + 0.044715 x^2)
    float u0, u1, t0, t1;
    unpack_f32x2(mul_f32x2(p, x), u0, u1);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
    const uint64_t hx = mul_f32x2(x, pack_f32x2(0.5f, 0.5f));
    return fma_f32x2(hx, pack_f32x2(t0, t1), hx);
}

SDOD_DEVICE float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
SDOD_DEVICE float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

}  // namespace sdod
