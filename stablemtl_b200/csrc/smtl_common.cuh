// Hand-written sm_100a primitives shared by every kernel of the StableMTL hot path:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld / st) and the
// shared-memory + instruction descriptors the 5th-gen tensor cores consume.
// No CUTLASS/CuTe is used; the bit layouts follow the PTX ISA "tcgen05" matrix/instruction
// descriptor tables.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace smtl {

#ifndef SMTL_SPIN_LIMIT
#define SMTL_SPIN_LIMIT (1u << 24)   // watchdog: a wedged mbarrier traps instead of hanging the GPU
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ----------------------------------------------------------------------------- explicit shared-space accesses
// The kernels align their dynamic shared memory by integer arithmetic on the pointer, after which the compiler no longer
// knows the address space and emits GENERIC loads / stores (LD.E / ST.E through the global pipeline: `lg` stalls in
// ncu) for plain dereferences.  Staging tiles and on-chip accumulators go through these instead.
__device__ __forceinline__ void sts_u16(uint32_t addr, uint16_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint16_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v2_s64(uint32_t addr, long long a, long long b) {
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void lds_v2_s64(uint32_t addr, long long& a, long long& b) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr) : "memory");
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > SMTL_SPIN_LIMIT) {
            printf("smtl: mbarrier watchdog block=(%d,%d) thread=%d bar=%u parity=%u\n", blockIdx.x, blockIdx.y,
                   threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}

// Long waits (an epilogue warp waiting for a whole tile's MMAs): back off between polls so that idle warps do not
// compete with the producer / MMA warps for the barrier unit.
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(ns);
        if (++spins > SMTL_SPIN_LIMIT) {
            printf("smtl: mbarrier watchdog block=(%d,%d) thread=%d bar=%u parity=%u\n", blockIdx.x, blockIdx.y,
                   threadIdx.x, smem_u32(bar), parity);
            __trap();
        }
    }
}

// Warp-collective wait: lane 0 polls, the other lanes park at the warp barrier.  An mbarrier wait executed by all 32
// lanes is 32 serialised barrier-unit operations (~430 clk per warp-wide wait measured on B200, and the spinning lanes
// of idle warps slow everyone else's barrier traffic); __syncwarp() orders the other lanes after lane 0's observation.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// 2-D tiled load global -> shared, completion on an mbarrier (coordinates are signed: OOB rows are zero-filled)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0,
                                            int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ----------------------------------------------------------------------------- clusters / CTA pairs
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) as seen in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// Arrive on an mbarrier of another CTA of the cluster.  Default semantics (release at CTA scope) -- NOT `.release.cluster`:
// that form compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive, i.e. the epilogue warp waited for every global
// store of its tile to be acknowledged before it handed the accumulator back (ncu: 4 % of all samples on the ERRBAR of a
// K = 1024 conv).  What the peer's MMA warp needs ordered is this warp's TMEM reads, and those are complete at
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load MULTICAST to the CTAs of `cta_mask`: the box lands at the same shared-memory offset in each of them and
// complete_tx is signalled on the mbarrier at the same offset in each of them (one L2 read feeds the whole cluster).
__device__ __forceinline__ void tma_load_2d_mcast(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int32_t c0, int32_t c1,
                                                  uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
        : "memory");
}
// commit all prior tcgen05.mma of this thread (cta_group::1); arrive(1) on the mbarrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void tc_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"(cta_mask)
        : "memory");
}
// TMA load issued by one CTA of a pair whose completion is signalled on an mbarrier that may live in the PEER CTA
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tmap, uint32_t bar_cluster_addr,
                                                 int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// commit all prior tcgen05.mma of this thread; arrive(1) on the mbarrier at the same offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs: 128 rows each] * B[smem: N/2 rows from each CTA]
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// commit all prior tcgen05.mma of this thread; arrive(1) on the mbarrier when they retire
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]   (kind::f16: bf16/fp16 inputs, fp32 accumulate)
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Shared-memory matrix descriptor for a K-major (or MN-major) bf16 tile written by TMA with 128-byte swizzle.
// Rows are 128 B (64 bf16); 8-row groups are 1024 B apart (SBO); LBO is unused for a single swizzle atom.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);   // start address, bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                       // leading byte offset (unused with swizzle) = 1
    d |= static_cast<uint64_t>(1024 >> 4) << 32;               // stride byte offset = 1024 B, bits [32,46)
    d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                       // layout type: SWIZZLE_128B
    return d;
}
// same, with an explicit leading byte offset (needed for MN-major operands wider than one 64-element atom)
__device__ __forceinline__ uint64_t make_smem_desc_sw128_lbo(uint32_t smem_addr, uint32_t lbo_bytes,
                                                             uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// 16-bit operand formats of the path (SMTL_FMT_* in the C header): both run at the same kind::f16 tensor rate.
constexpr int FMT_BF16 = 0;
constexpr int FMT_F16 = 1;

// Instruction descriptor, kind::f16, 16-bit x 16-bit -> fp32. a_mn / b_mn = 1 selects an MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc_16(int M, int N, int a_mn, int b_mn, int fmt) {
    return (1u << 4)                                  // D format: F32
           | ((fmt == FMT_F16 ? 0u : 1u) << 7)        // A format: F16 = 0, BF16 = 1
           | ((fmt == FMT_F16 ? 0u : 1u) << 10)       // B format
           | (static_cast<uint32_t>(a_mn) << 15)      // A major
           | (static_cast<uint32_t>(b_mn) << 16)      // B major
           | (static_cast<uint32_t>(N >> 3) << 17)    // N / 8
           | (static_cast<uint32_t>(M >> 4) << 24);   // M / 16
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (thread i gets lane i's row)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// TMEM -> registers in the MMA-fragment shape: 16 lanes x (4 x 256 bits = 32 fp32 columns).  Thread t = (g = t / 4, q = t % 4)
// gets r[4k + 2j + e] = (lane lane0 + g + 8j, column col0 + 8k + 2q + e) -- lanes g and g + 8, column pairs, like the
// accumulator fragment of mma.m16n8.  A packed pair {r[4k+2j], r[4k+2j+1]} is then one stmatrix / ldmatrix fragment.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// Four 8x8 16-bit matrices, registers <-> shared memory, TRANSPOSED on the way: thread 8i + r supplies the address of the
// 16-byte row r of matrix i; a thread's register of matrix i holds elements (2q, g) and (2q + 1, g) of the stored rows x
// columns, i.e. with fragments indexed [g][2q + e] the stored matrix is the transpose.
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("stmatrix.sync.aligned.x4.trans.m8n8.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c),
                 "r"(d) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                 : "r"(addr) : "memory");
}

// registers -> TMEM, same shape
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------- small math helpers
// packed fp32 pairs: Blackwell issues two fp32 lanes per slot (FFMA2 / FADD2); results are those of the scalar ops
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
// erf-GELU for a PAIR on the FMA pipe alone (no SFU): gelu(x) = x * Phi(x), Phi(x) = 0.5 + xc * R(xc^2) with xc = x
// clamped to +-4.5 (Phi(-4.5) = 3.4e-6) and R a degree-10 polynomial in t = 2 xc^2 / 4.5^2 - 1 (Chebyshev fit of
// (Phi(x) - 0.5) / x, converted to the monomial basis in t: well conditioned in fp32).  |error| < 3e-6 for |x| <= 4.5 and
// < 3e-6 * |x| beyond, far below the 16-bit rounding of the value.  Packed FFMA2: ~9 issue slots per element against
// 15 + 2 SFU ops (rcp, ex2) for the Abramowitz-Stegun erf it replaced -- the GEGLU / per-task MLP epilogues are
// issue- and SFU-bound at K = 320.
__device__ __forceinline__ float2 gelu_poly2(float2 x) {
    float2 xc;
    xc.x = fminf(fmaxf(x.x, -4.5f), 4.5f);
    xc.y = fminf(fmaxf(x.y, -4.5f), 4.5f);
    const float2 s = ffma2(xc, xc, make_float2(0.f, 0.f));
    const float k = 2.0f / (4.5f * 4.5f);
    const float2 t = ffma2(s, make_float2(k, k), make_float2(-1.0f, -1.0f));
#define SMTL_C2(v) make_float2(v, v)
    float2 r = ffma2(t, SMTL_C2(0.000789802405051887f), SMTL_C2(-0.002244635485112667f));
    r = ffma2(r, t, SMTL_C2(0.00314583582803607f));
    r = ffma2(r, t, SMTL_C2(-0.0054423320107162f));
    r = ffma2(r, t, SMTL_C2(0.01110508106648922f));
    r = ffma2(r, t, SMTL_C2(-0.018895722925662994f));
    r = ffma2(r, t, SMTL_C2(0.028388366103172302f));
    r = ffma2(r, t, SMTL_C2(-0.040139030665159225f));
    r = ffma2(r, t, SMTL_C2(0.05468999966979027f));
    r = ffma2(r, t, SMTL_C2(-0.07719199359416962f));
    r = ffma2(r, t, SMTL_C2(0.15690511465072632f));
    const float2 ph = ffma2(xc, r, SMTL_C2(0.5f));
#undef SMTL_C2
    return ffma2(x, ph, make_float2(0.f, 0.f));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
// x * sigmoid(x) with sigmoid(x) = 0.5 tanh(0.5 x) + 0.5: ONE SFU op (tanh.approx, rel. error ~2^-11, below the
// 16-bit rounding of the value it feeds) instead of ex2 + a full-precision division
__device__ __forceinline__ float silu_fast(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return x * fmaf(t, 0.5f, 0.5f);
}
// x * sigmoid(x) to ~2 ulp: ex2.approx + rcp.approx (two SFU ops)
__device__ __forceinline__ float silu_exact(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
// exact unsigned division by a runtime constant: q = (n * mul) >> 32 >> shr  (n < 2^31, d >= 1)
struct FastDiv {
    uint32_t mul, shr, d;
};
__host__ inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    f.d = d;
    if (d == 1) { f.mul = 0; f.shr = 0; return f; }
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;                    // ceil(log2 d)
    f.mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    f.shr = l - 1;
    return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
    if (f.d == 1) return n;
    const uint32_t t = __umulhi(n, f.mul);
    return (t + ((n - t) >> 1)) >> f.shr;
}

// fp32 -> 16-bit pair.  fp16 conversions saturate at +-65504 instead of overflowing to inf: ONE instruction
// (F2FP.SATFINITE.F16.F32.PACK_AB) -- the clamp used to be four FMNMX per pair in every epilogue.
__device__ __forceinline__ uint32_t pack16x2(float lo, float hi, int fmt) {
    if (fmt == FMT_F16) {
        uint32_t r;
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    }
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
// same without the +-65504 clamp, for values known to be in range (softmax probabilities)
__device__ __forceinline__ uint32_t pack16x2_nosat(float lo, float hi, int fmt) {
    if (fmt == FMT_F16) {
        __half2 v = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack16x2(uint32_t u, int fmt) {
    if (fmt == FMT_F16) {
        __half2 v = *reinterpret_cast<__half2*>(&u);
        return __half22float2(v);
    }
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}
__device__ __forceinline__ uint16_t to16(float x, int fmt) {
    return static_cast<uint16_t>(pack16x2(x, 0.f, fmt) & 0xFFFFu);
}

// ----------------------------------------------------------------------------- deterministic channel statistics
// A GroupNorm statistics cell is FOUR int64 fixed-point accumulators (sum_lo, sum_hi, sq_lo, sq_hi): a partial sum p goes,
// exactly, into the fine accumulator (scale 2^32, |p| < 2^12) or the coarse one (scale 2^8).  Integer addition is
// associative, so the atomics of the producing GEMM's CTAs give the same bits in whatever order they land (fp32
// atomicAdd did not: run-to-run differences of ~1e-7 that the network amplified to the parity tolerance).
// Range: a cell receives at most rows_per_image / 32 partials (9600 at 480x640, full resolution); fine partials are
// < 2^44 and coarse ones (32 rows of x^2, |x| <= 65504) < 2^45, so a cell stays below 2^63 up to 2^18 partials.
// Consumers convert each CELL to double before they add cells up (a group's total need not fit an int64).
constexpr float STATS_LO_LIMIT = 4096.0f;            // 2^12
constexpr float STATS_LO_SCALE = 4294967296.0f;      // 2^32
constexpr float STATS_HI_SCALE = 256.0f;             // 2^8
__device__ __forceinline__ long long stats_fix(float p, bool& hi) {
    hi = !(fabsf(p) < STATS_LO_LIMIT);
    return __float2ll_rn(p * (hi ? STATS_HI_SCALE : STATS_LO_SCALE));
}
__device__ __forceinline__ void stats_atomic_add(unsigned long long* cell, float s, float q) {
    bool hi;
    long long v = stats_fix(s, hi);
    atomicAdd(cell + (hi ? 1 : 0), (unsigned long long)v);
    v = stats_fix(q, hi);
    atomicAdd(cell + (hi ? 3 : 2), (unsigned long long)v);
}
__device__ __forceinline__ double stats_value(long long lo, long long hi) {
    return (double)lo * (1.0 / 4294967296.0) + (double)hi * (1.0 / 256.0);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Lane j <- sum over the 32 lanes of s[j] (recursive halving: 31 shuffles instead of 32 x 5).
__device__ __forceinline__ float warp_transpose_sum(float (&s)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float mine = upper ? s[i + off] : s[i];
            const float other = upper ? s[i] : s[i + off];
            s[i] = mine + __shfl_xor_sync(0xffffffffu, other, off);
        }
    }
    return s[0];
}

// N values per lane (N = 1, 2, ... 32, a power of two): lane L <- the warp total of value L / (32 / N).  The first
// log2(N) stages halve the number of values (splitting on lane bits 4, 3, ...), the rest are a plain butterfly:
// N - 1 + 5 - log2(N) shuffles.  N = 32 is warp_transpose_sum.
template <int N>
__device__ __forceinline__ float warp_transpose_sum_n(float (&s)[N], int lane) {
    int off = 16;
#pragma unroll
    for (int n = N / 2; n >= 1; n >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const float mine = upper ? s[i + n] : s[i];
            const float other = upper ? s[i] : s[i + n];
            s[i] = mine + __shfl_xor_sync(0xffffffffu, other, off);
        }
        off >>= 1;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        if (off >= 1) s[0] += __shfl_xor_sync(0xffffffffu, s[0], off);
        off >>= 1;
    }
    return s[0];
}
// Column statistics of one 32-row x 32-column epilogue chunk (lane = row, v[j] = column j), G adjacent columns summed
// in the thread first: lane L with L % G == 0 gets the sum / sum of squares of columns L .. L + G - 1 over the 32 rows.
template <int G>
__device__ __forceinline__ void chunk_col_stats(const float (&v)[32], bool row_ok, int lane, float& cs, float& cq) {
    constexpr int N = 32 / G;
    float s[N], q[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const float x = row_ok ? v[i * G + k] : 0.0f;
            a += x;
            b = fmaf(x, x, b);
        }
        s[i] = a;
        q[i] = b;
    }
    cs = warp_transpose_sum_n<N>(s, lane);
    cq = warp_transpose_sum_n<N>(q, lane);
}

}  // namespace smtl
