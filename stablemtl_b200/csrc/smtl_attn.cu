// smtl_attn.cu -- flash-style self-attention for the SD-2 UNet transformer blocks (head dim 64).
// Replaces xformers.ops.memory_efficient_attention at reference src/model/attention.py:391-397,417.
//
// The kernels (smtl_fattn4_kernel for the UNet, smtl_vattn_kernel for the VAE mid-block) are described where they are defined.
#include "smtl_common.cuh"
#include "smtl_host.h"

namespace {
using namespace smtl;

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE_BYTES = 128 * 64 * 2;       // 16 KB: a [128 x 64] bf16 tile

struct alignas(64) FattnKParams {
    CUtensorMap tm;
    CUtensorMap tm_half;                 // same tensor, [64 x 64] box (smtl_vattn_kernel: each CTA of a pair loads half a tile)
    int32_t q_col0, k_col0, v_col0;
    int32_t ntok, heads;
    uint16_t* out;
    int32_t ldo;
    float scale_log2;
    int32_t fmt;
};

// ================================================================================================ shared pieces
// Both attention kernels of this file run the same pipeline:
//   warp 0      : TMA producer  (Q once; K/V tiles through a ring)
//   warp 1      : MMA issuer    S = Q K^T, then O += P V with P read straight from TMEM (tcgen05.mma with the A operand
//                               in tensor memory)
//   softmax     : one query row per thread: tcgen05.ld S -> running max (lazy: the reference max only moves when it
//                 grows by > 2^8) -> p = ex2(s*scale - m) -> 16-bit pairs -> tcgen05.st over the S columns just consumed
constexpr float F2_LAZY = 8.0f;                            // log2 units

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// (A polynomial ex2 on the FMA pipe for 1/8 .. 3/8 of the scores -- the FlashAttention-4 trick, on packed FFMA2 / FADD2.RM --
// was measured in both generations of the d = 64 kernel and lost 3-12 % each time: 10 issue slots per pair against 2
// MUFU.EX2, and the softmax warps are short of issue slots before they are short of SFU cycles.)
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

// (Measured and dropped in round 2, on the first d = 64 kernel -- two query tiles per CTA, one softmax warp of each per SM
// sub-partition: passing an "exp token" over named barriers between those two warps so that their exponential phases run
// back to back.  0.850 ms before and after; what fixed that kernel's lock-step was more, shorter streams: smtl_fattn4_kernel.)
template <bool MASKED>
__device__ __forceinline__ void softmax_tile(const FattnKParams& p, uint32_t t_s, int kv_valid, float& m_run,
                                             float& l_run, uint32_t t_o, bool have_o, int o_chunks = 2,
                                             uint64_t* pv_bar = nullptr, uint32_t pv_parity = 0) {
    // the whole S row (128 fp32) is pulled into registers with four back-to-back tcgen05.ld and ONE wait
    uint32_t rr[128];
#pragma unroll
    for (int c = 0; c < 4; ++c) tmem_ld_32x32(t_s + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&rr[c * 32]));
    tmem_ld_wait();
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 128; i += 4) {
        float a = __uint_as_float(rr[i]), b = __uint_as_float(rr[i + 1]), c = __uint_as_float(rr[i + 2]),
              d = __uint_as_float(rr[i + 3]);
        if (MASKED) {
            a = (i < kv_valid) ? a : -INFINITY;
            b = (i + 1 < kv_valid) ? b : -INFINITY;
            c = (i + 2 < kv_valid) ? c : -INFINITY;
            d = (i + 3 < kv_valid) ? d : -INFINITY;
        }
        mx0 = fmaxf(mx0, a); mx1 = fmaxf(mx1, b); mx2 = fmaxf(mx2, c); mx3 = fmaxf(mx3, d);
    }
    const float mx_s = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
    // lazy running max: exact as long as the SAME reference max is used for p, l and O
    const bool grow = mx_s > m_run + F2_LAZY;
    if (__any_sync(0xffffffffu, grow)) {
        const float alpha = grow ? ex2_approx(m_run - mx_s) : 1.0f;     // 0 on the first tile (m_run = -inf)
        if (grow) { m_run = mx_s; l_run *= alpha; }
        if (have_o) {
            if (pv_bar) {                                     // smtl_vattn_kernel issues S(j) BEFORE PV(j-1): O must hold
                mbar_wait(pv_bar, pv_parity);                 // every earlier tile before it is rescaled
                tc_fence_after();
            }
#pragma unroll 1
            for (int c = 0; c < o_chunks; ++c) {              // rare path: keep its register footprint small
                uint32_t oo[32];
                tmem_ld_32x32(t_o + c * 32, oo);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) oo[i] = __float_as_uint(__uint_as_float(oo[i]) * alpha);
                tmem_st_32x32(t_o + c * 32, oo);
            }
        }
    }
    const float neg_m = -m_run;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            float p0 = ex2_approx(fmaf(__uint_as_float(rr[c * 32 + i]), p.scale_log2, neg_m));
            float p1 = ex2_approx(fmaf(__uint_as_float(rr[c * 32 + i + 1]), p.scale_log2, neg_m));
            if (MASKED) {
                p0 = (c * 32 + i < kv_valid) ? p0 : 0.f;
                p1 = (c * 32 + i + 1 < kv_valid) ? p1 : 0.f;
            }
            ps0 += p0;
            ps1 += p1;
            pk[i >> 1] = pack16x2_nosat(p0, p1, p.fmt);    // p <= 2^8: no fp16 saturation needed
        }
        tmem_st_32x16(t_s + c * 16, pk);          // P aliases the S columns (all of S is in registers by now)
    }
    l_run += ps0 + ps1;
    tmem_st_wait();
}

// ================================================================================================ d = 64, four query tiles
// smtl_fattn4_kernel: the same pipeline with FOUR 128-row query tiles per CTA and 64-key steps (TMEM: S/P 4 x 64 columns,
// O 4 x 64).  Why: in the two-tile kernel an SM sub-partition hosts two softmax warps that run IN PHASE -- both in their
// exponential phase (sharing the SFU), then both waiting on the tensor pipe's P·V + next Q·K^T round trip -- so a tile pair
// costs SFU time PLUS round-trip time (ncu: SFU 58 %, tensor 29 %, issue 41 % busy; 3500 clk per pair of tiles against
// 2048 SFU clk).  Four independent streams per sub-partition, each with half the work per step, overlap one stream's
// round trip with the others' exponentials.  K and V are still staged as 128-key boxes (one TMA box per operand and
// stage); a step consumes half a box.  The softmax arithmetic is packed (FFMA2 / FADD2: two fp32 lanes per issue slot).
constexpr int F4_NQ = 4;
constexpr int F4_KS = 64;                                  // keys per step
constexpr int F4_THREADS = 64 + F4_NQ * 128;               // 576: TMA, MMA, 16 softmax warps
// Five warps on two of the SM sub-partitions: 96 registers per thread (set with __maxnreg__; __launch_bounds__ would cap the
// kernel at 88 and spill 130 words of the softmax row).
constexpr int F4_REGS = 96;
constexpr int F4_NS = 4;                                   // K/V ring stages (one 128-key box of K and of V: 32 KB)
constexpr int F4_SMEM = F4_NQ * TILE_BYTES + F4_NS * 2 * TILE_BYTES + 256;

template <bool MASKED, int FMT>
__device__ __forceinline__ void softmax_step64(const FattnKParams& p, uint32_t t_s, int kv_valid, float& m_run,
                                               float& l_run, uint32_t t_o, bool have_o) {
    uint32_t rr[64];
    tmem_ld_32x32(t_s, *reinterpret_cast<uint32_t(*)[32]>(&rr[0]));
    tmem_ld_32x32(t_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&rr[32]));
    tmem_ld_wait();
    float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
        float a = __uint_as_float(rr[i]), b = __uint_as_float(rr[i + 1]), c = __uint_as_float(rr[i + 2]),
              d = __uint_as_float(rr[i + 3]);
        if (MASKED) {
            a = (i < kv_valid) ? a : -INFINITY;
            b = (i + 1 < kv_valid) ? b : -INFINITY;
            c = (i + 2 < kv_valid) ? c : -INFINITY;
            d = (i + 3 < kv_valid) ? d : -INFINITY;
        }
        mx0 = fmaxf(mx0, a); mx1 = fmaxf(mx1, b); mx2 = fmaxf(mx2, c); mx3 = fmaxf(mx3, d);
    }
    const float mx_s = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;
    const bool grow = mx_s > m_run + F2_LAZY;                 // lazy running max, as in softmax_tile
    if (__any_sync(0xffffffffu, grow)) {
        const float alpha = grow ? ex2_approx(m_run - mx_s) : 1.0f;
        if (grow) { m_run = mx_s; l_run *= alpha; }
        if (have_o) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {                     // rare path, 16 columns at a time: the S row stays live
                uint32_t oo[16];
                tmem_ld_32x16(t_o + c * 16, oo);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) oo[i] = __float_as_uint(__uint_as_float(oo[i]) * alpha);
                tmem_st_32x16(t_o + c * 16, oo);
            }
        }
    }
    const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(-m_run, -m_run);
    float2 ps = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
            const float2 t = ffma2(make_float2(__uint_as_float(rr[c * 32 + i]), __uint_as_float(rr[c * 32 + i + 1])), sc2, nm2);
            float2 e = make_float2(ex2_approx(t.x), ex2_approx(t.y));
            if (MASKED) {
                e.x = (c * 32 + i < kv_valid) ? e.x : 0.f;
                e.y = (c * 32 + i + 1 < kv_valid) ? e.y : 0.f;
            }
            ps = fadd2(ps, e);
            pk[i >> 1] = pack16x2_nosat(e.x, e.y, FMT);     // p <= 2^8: no fp16 saturation needed
        }
        tmem_st_32x16(t_s + c * 16, pk);          // P aliases the S columns (all of S is in registers by now)
    }
    l_run += ps.x + ps.y;
    tmem_st_wait();
}

template <int FMT>
__global__ void __maxnreg__(F4_REGS) smtl_fattn4_kernel(const __grid_constant__ FattnKParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;                                    // 4 tiles
    uint8_t* sK = smem + F4_NQ * TILE_BYTES;               // NS boxes
    uint8_t* sV = sK + F4_NS * TILE_BYTES;                 // NS boxes
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + F4_NS * TILE_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;                // [NS]
    uint64_t* kv_empty = bars + 1 + F4_NS;       // [NS]
    uint64_t* s_full = bars + 1 + 2 * F4_NS;     // [NQ]
    uint64_t* p_full = s_full + F4_NQ;           // [NQ]
    uint64_t* o_done = p_full + F4_NQ;           // [NQ]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + F4_NQ);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * F4_NQ * BQ;
    const int b = blockIdx.y / p.heads;
    const int hd = blockIdx.y - b * p.heads;
    const int row_base = b * p.ntok;
    const int nbox = (p.ntok + BKV - 1) / BKV;             // staged 128-key boxes
    const int nsub = (p.ntok + F4_KS - 1) / F4_KS;         // 64-key steps
    const int ntile_q = min(F4_NQ, (p.ntok - q0 + BQ - 1) / BQ);

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) { printf("smtl_fattn4: smem base not 1024-aligned\n"); __trap(); }
        tma_prefetch_desc(&p.tm);
        mbar_init(q_full, 1);
        for (int s = 0; s < F4_NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int w = 0; w < F4_NQ; ++w) { mbar_init(&s_full[w], 1); mbar_init(&p_full[w], 4); mbar_init(&o_done[w], 1); }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, ntile_q * TILE_BYTES);
            for (int w = 0; w < ntile_q; ++w)
                tma_load_2d(sQ + w * TILE_BYTES, &p.tm, q_full, p.q_col0 + hd * HD, row_base + q0 + w * BQ);
        }
        __syncwarp();
        for (int j = 0; j < nbox; ++j) {
            const int s = j % F4_NS;
            mbar_wait(&kv_empty[s], ((j / F4_NS) & 1) ^ 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
                tma_load_2d(sK + s * TILE_BYTES, &p.tm, &kv_full[s], p.k_col0 + hd * HD, row_base + j * BKV);
                tma_load_2d(sV + s * TILE_BYTES, &p.tm, &kv_full[s], p.v_col0 + hd * HD, row_base + j * BKV);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        const uint32_t IDESC_S = make_idesc_16(BQ, F4_KS, 0, 0, FMT);  // S[128,64] = Q[128,64] K[64,64]^T
        const uint32_t IDESC_O = make_idesc_16(BQ, HD, 0, 1, FMT);       // O[128,64] += P[128,64] V[64,64] (V MN-major)
        auto issue_s = [&](int w, int j) {                               // step j: box j / 2, half j % 2
            const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ + w * TILE_BYTES));
            const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + ((j >> 1) % F4_NS) * TILE_BYTES) + (j & 1) * F4_KS * 128);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) tc_mma_f16(tmem_base + w * F4_KS, dq + 2 * k, dk + 2 * k, IDESC_S, k != 0);
            tc_commit(&s_full[w]);
        };
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        if (elect_one())
            for (int w = 0; w < ntile_q; ++w) issue_s(w, 0);
        __syncwarp();
        for (int j = 0; j < nsub; ++j) {
            const int s = (j >> 1) % F4_NS;
            for (int w = 0; w < ntile_q; ++w) {
                mbar_wait(&p_full[w], j & 1);                 // softmax wrote P_w(j) and rescaled O_w
                if (w == 0 && j + 1 < nsub && ((j + 1) & 1) == 0) {      // S(j+1) opens the next box
                    const int nb = (j + 1) >> 1;
                    mbar_wait(&kv_full[nb % F4_NS], (nb / F4_NS) & 1);
                }
                tc_fence_after();
                const uint32_t sv = smem_u32(sV + s * TILE_BYTES) + (j & 1) * F4_KS * 128;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < F4_KS / 16; ++k) {
                        // V rows (kv) 16k .. 16k+16 of this half box; P: 8 TMEM columns per K step
                        const uint64_t dv = make_smem_desc_sw128(sv + k * 16 * 128);
                        tc_mma_f16_ts(tmem_base + 256 + w * HD, tmem_base + w * F4_KS + 8 * k, dv, IDESC_O, (j | k) != 0);
                    }
                    if (j + 1 < nsub) issue_s(w, j + 1);      // in-order after PV_w(j): may overwrite P_w(j)
                    else tc_commit(&o_done[w]);
                }
                __syncwarp();
            }
            if ((j & 1) || j == nsub - 1) {                   // both halves of the box consumed by every MMA issued so far
                if (elect_one()) tc_commit(&kv_empty[s]);
                __syncwarp();
            }
        }
    } else {
        const int w = (warp - 2) >> 2;                         // which query tile
        if (w < ntile_q) {
            const int quarter = warp & 3;                      // TMEM lane quarter this warp may access
            const int r = quarter * 32 + lane;                 // query row within the tile == TMEM lane
            const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
            const uint32_t t_s = tmem_base + w * F4_KS + lane_off;
            const uint32_t t_o = tmem_base + 256 + w * HD + lane_off;
            float m_run = -INFINITY, l_run = 0.f;
            const int tail = p.ntok - (nsub - 1) * F4_KS;      // valid kv columns of the last step
            for (int j = 0; j < nsub; ++j) {
                mbar_wait(&s_full[w], j & 1);                  // S_w(j) ready; implies PV_w(j-1) retired
                tc_fence_after();
                if (j == nsub - 1 && tail < F4_KS) softmax_step64<true, FMT>(p, t_s, tail, m_run, l_run, t_o, j > 0);
                else softmax_step64<false, FMT>(p, t_s, F4_KS, m_run, l_run, t_o, j > 0);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[w]);
            }
            mbar_wait(&o_done[w], 0);
            tc_fence_after();
            const float inv = 1.0f / l_run;
            const int qrow = q0 + w * BQ + r;
            const bool row_ok = qrow < p.ntok;
            uint16_t* dst = p.out + (int64_t)(row_base + qrow) * p.ldo + hd * HD;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t rr[32];
                tmem_ld_32x32(t_o + c * 32, rr);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 o;
                        o.x = pack16x2(__uint_as_float(rr[i]) * inv, __uint_as_float(rr[i + 1]) * inv, FMT);
                        o.y = pack16x2(__uint_as_float(rr[i + 2]) * inv, __uint_as_float(rr[i + 3]) * inv, FMT);
                        o.z = pack16x2(__uint_as_float(rr[i + 4]) * inv, __uint_as_float(rr[i + 5]) * inv, FMT);
                        o.w = pack16x2(__uint_as_float(rr[i + 6]) * inv, __uint_as_float(rr[i + 7]) * inv, FMT);
                        *reinterpret_cast<uint4*>(dst + c * 32 + i) = o;
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// ================================================================================================ d = 512, one head
// The VAE mid-block attention (diffusers Attention(heads = 1, dim_head = 512) inside UNetMidBlock2D, reached from
// src/stablemtl_pipeline.py:619-620,642-643): N = H/8 * W/8 tokens per image, ONE head of 512
// channels.  Round 1 ran it per image as QK^T -> fp32 S in HBM (92 MB at 480x640) -> row softmax -> PV: four launches
// per image, 5 % of the step for 1.7 % of its flops.  Here S and P never leave the SM:
//   one CTA per 128 query rows of one image; Q (128 x 512) stays in shared memory as eight 64-wide slices;
//   warp 0      : TMA producer -- K and V stream through a ring of [128 x 64] tiles (8 K slices, 4 V slices per KV tile)
//   warp 1      : MMA issuer   -- S(j) = sum over the 8 slices of Q_s K_s^T into one of TWO S buffers in TMEM, so the
//                                 tensor pipe works on S(j+1) while the softmax warps work on S(j);
//                                 O += P(j) V(j) with P read from TMEM (it overwrites the S columns it came from)
//   warps 2..5  : softmax      -- one query row per thread (lazy running max, ex2.approx)
// TMEM holds 512 fp32 columns: two S buffers (256) leave room for HALF of O (256 of its 512 columns), so the kernel
// makes two passes over the keys, one per half of the value channels, recomputing S and the softmax (1.5x the
// algorithmic flops; the pass is tensor-bound at 3072 clk per KV tile against ~1500 clk of softmax).
// TMEM columns: S0/P0 [0,128)  S1/P1 [128,256)  O half [256,512).
//
// A CTA consumes 192 KB of K / V per 3072 tensor cycles -- 62 B/clk per SM, 16 TB/s over 148 SMs at full clock, about
// twice what L2 delivers: the MMA warp spent 42 % of its time waiting for a full ring stage (ncu: tensor pipe 48.6 %
// active).  So two CTAs on neighbouring query tiles of the same image form a CLUSTER: each loads HALF of every K / V
// tile (64 of its 128 rows) and the TMA multicasts it into both CTAs' rings -- one L2 read per pair.  A stage is free
// again when BOTH CTAs' MMAs have retired it: tcgen05.commit multicasts its arrive to both CTAs' empty barriers.
constexpr int V5_THREADS = 192;
constexpr int V5_SLICES = 8;                               // 512 / 64
constexpr int V5_NS = 6;                                   // ring stages (16 KB each)
constexpr int V5_SMEM = V5_SLICES * TILE_BYTES + V5_NS * TILE_BYTES + 256;

__global__ void __launch_bounds__(V5_THREADS, 1) smtl_vattn_kernel(const __grid_constant__ FattnKParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;                                    // 8 slices [128 x 64]
    uint8_t* sR = smem + V5_SLICES * TILE_BYTES;           // K / V ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(sR + V5_NS * TILE_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;                // [NS]
    uint64_t* kv_empty = bars + 1 + V5_NS;       // [NS]
    uint64_t* s_full = bars + 1 + 2 * V5_NS;     // [2]   MMA -> softmax
    uint64_t* p_full = s_full + 2;               // [2]   softmax -> MMA
    uint64_t* o_done = p_full + 2;               //       MMA -> epilogue, once per pass
    uint64_t* o_free = o_done + 1;               //       epilogue -> MMA, once per pass
    uint64_t* pv_done = o_free + 1;              //       MMA -> softmax: PV(j) retired (needed before O is rescaled)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ;                        // may lie past the image (odd tile count): nothing is stored then
    const int row_base = blockIdx.y * p.ntok;
    const int ntiles = (p.ntok + BKV - 1) / BKV;
    const uint32_t rank = cluster_ctarank();               // 0 / 1 within the pair (cluster dims 2 x 1 x 1)

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) { printf("smtl_vattn: smem base not 1024-aligned\n"); __trap(); }
        tma_prefetch_desc(&p.tm);
        tma_prefetch_desc(&p.tm_half);
        mbar_init(q_full, 1);
        for (int s = 0; s < V5_NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 2); }   // empty: both CTAs' MMAs
        for (int w = 0; w < 2; ++w) { mbar_init(&s_full[w], 1); mbar_init(&p_full[w], 4); }
        mbar_init(o_done, 1);
        mbar_init(o_free, 4);
        mbar_init(pv_done, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    cluster_sync_all();                                    // both CTAs' barriers are initialised before either signals the other
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer (same order as the MMA warp consumes)
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, V5_SLICES * TILE_BYTES);
            for (int s = 0; s < V5_SLICES; ++s) tma_load_2d(sQ + s * TILE_BYTES, &p.tm, q_full, p.q_col0 + s * 64, row_base + q0);
        }
        __syncwarp();
        int st = 0;
        uint32_t ph = 0;
        auto load = [&](int col, int row) {
            mbar_wait(&kv_empty[st], ph ^ 1u);             // both CTAs have retired this stage
            if (elect_one()) {
                mbar_arrive_expect_tx(&kv_full[st], TILE_BYTES);               // my half + the peer's half land here
                tma_load_2d_mcast(sR + st * TILE_BYTES + rank * (TILE_BYTES / 2), &p.tm_half, &kv_full[st], col,
                                  row + (int)rank * (BKV / 2), (uint16_t)3);
            }
            __syncwarp();
            if (++st == V5_NS) { st = 0; ph ^= 1u; }
        };
        for (int h = 0; h < 2; ++h) {
            auto load_k = [&](int j) { for (int s = 0; s < V5_SLICES; ++s) load(p.k_col0 + s * 64, row_base + j * BKV); };
            auto load_v = [&](int j) { for (int s = 0; s < 4; ++s) load(p.v_col0 + h * 256 + s * 64, row_base + j * BKV); };
            load_k(0);
            if (ntiles > 1) load_k(1);
            for (int j = 0; j < ntiles; ++j) {
                load_v(j);
                if (j + 2 < ntiles) load_k(j + 2);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer (converged warp, elect.sync at the issue)
        const uint32_t IDESC_S = make_idesc_16(BQ, BKV, 0, 0, p.fmt);   // S[128,128] += Q_s[128,64] K_s[128,64]^T
        const uint32_t IDESC_O = make_idesc_16(BQ, 64, 0, 1, p.fmt);    // O[128,64]  += P[128,128] V_s[128,64]   (V MN-major)
        int st = 0;
        uint32_t ph = 0;
        uint32_t p_uses[2] = {0, 0};
        auto issue_s = [&](int j) {
            const int sb = j & 1;
            for (int s = 0; s < V5_SLICES; ++s) {
                mbar_wait(&kv_full[st], ph);
                if (elect_one()) {
                    const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ + s * TILE_BYTES));
                    const uint64_t dk = make_smem_desc_sw128(smem_u32(sR + st * TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k) tc_mma_f16(tmem_base + sb * 128, dq + 2 * k, dk + 2 * k, IDESC_S, (uint32_t)(s | k));
                    tc_commit_mcast(&kv_empty[st], (uint16_t)3);
                    if (s == V5_SLICES - 1) tc_commit(&s_full[sb]);
                }
                __syncwarp();
                if (++st == V5_NS) { st = 0; ph ^= 1u; }
            }
        };
        auto issue_pv = [&](int j) {
            const int sb = j & 1;
            for (int s = 0; s < 4; ++s) {
                mbar_wait(&kv_full[st], ph);
                if (elect_one()) {
                    const uint32_t sv = smem_u32(sR + st * TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BKV / 16; ++k) {
                        const uint64_t dv = make_smem_desc_sw128(sv + k * 16 * 128);       // kv rows 16k .. 16k + 16
                        tc_mma_f16_ts(tmem_base + 256 + s * 64, tmem_base + sb * 128 + 8 * k, dv, IDESC_O, (uint32_t)(j | k));
                    }
                    tc_commit_mcast(&kv_empty[st], (uint16_t)3);
                    if (s == 3) tc_commit(pv_done);
                }
                __syncwarp();
                if (++st == V5_NS) { st = 0; ph ^= 1u; }
            }
        };
        mbar_wait(q_full, 0);
        for (int h = 0; h < 2; ++h) {
            issue_s(0);
            if (ntiles > 1) issue_s(1);
            for (int j = 0; j < ntiles; ++j) {
                const int sb = j & 1;
                mbar_wait(&p_full[sb], p_uses[sb] & 1u);       // softmax wrote P(j) over S(j) and rescaled O
                ++p_uses[sb];
                if (j == 0 && h > 0) mbar_wait(o_free, (uint32_t)((h - 1) & 1));   // the epilogue has read the previous half of O
                tc_fence_after();
                issue_pv(j);
                if (j + 2 < ntiles) issue_s(j + 2);           // in order after PV(j): may overwrite P(j)
            }
            if (elect_one()) tc_commit(o_done);
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------- softmax + epilogue: one query row per thread
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        const uint32_t t_o = tmem_base + 256 + lane_off;
        const int tail = p.ntok - (ntiles - 1) * BKV;
        const int qrow = q0 + r;
        const bool row_ok = qrow < p.ntok;
        uint32_t s_uses[2] = {0, 0};
        for (int h = 0; h < 2; ++h) {
            float m_run = -INFINITY, l_run = 0.f;
            for (int j = 0; j < ntiles; ++j) {
                const int sb = j & 1;
                mbar_wait(&s_full[sb], s_uses[sb] & 1u);
                ++s_uses[sb];
                tc_fence_after();
                const uint32_t t_s = tmem_base + sb * 128 + lane_off;
                const uint32_t pv_par = (uint32_t)((h * ntiles + j - 1) & 1);      // phase in which PV(j-1) completes
                if (j == ntiles - 1 && tail < BKV) softmax_tile<true>(p, t_s, tail, m_run, l_run, t_o, j > 0, 8, pv_done, pv_par);
                else softmax_tile<false>(p, t_s, BKV, m_run, l_run, t_o, j > 0, 8, pv_done, pv_par);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[sb]);
            }
            mbar_wait(o_done, (uint32_t)(h & 1));
            tc_fence_after();
            const float inv = 1.0f / l_run;
            uint16_t* dst = p.out + (int64_t)(row_base + qrow) * p.ldo + h * 256;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
                uint32_t rr[32];
                tmem_ld_32x32(t_o + c * 32, rr);
                tmem_ld_wait();
                if (row_ok) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        uint4 o;
                        o.x = pack16x2(__uint_as_float(rr[i]) * inv, __uint_as_float(rr[i + 1]) * inv, p.fmt);
                        o.y = pack16x2(__uint_as_float(rr[i + 2]) * inv, __uint_as_float(rr[i + 3]) * inv, p.fmt);
                        o.z = pack16x2(__uint_as_float(rr[i + 4]) * inv, __uint_as_float(rr[i + 5]) * inv, p.fmt);
                        o.w = pack16x2(__uint_as_float(rr[i + 6]) * inv, __uint_as_float(rr[i + 7]) * inv, p.fmt);
                        *reinterpret_cast<uint4*>(dst + c * 32 + i) = o;
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_free);
        }
    }

    tc_fence_before();
    cluster_sync_all();                                    // the peer may still be signalling this CTA's barriers
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

extern "C" int smtl_fattn_plan(const smtl_fattn_args* a, smtl_fattn_op* op) {
    SMTL_CHECK_ARG(a && op && a->qkv && a->out_bf16, "fattn_plan: NULL argument");
    SMTL_CHECK_ARG(a->batch > 0 && a->ntok > 0 && a->heads > 0, "fattn_plan: empty problem");
    SMTL_CHECK_ARG(a->ld % 8 == 0 && a->ldo % 8 == 0 && a->q_col0 % 8 == 0 && a->k_col0 % 8 == 0 && a->v_col0 % 8 == 0,
                   "fattn_plan: unaligned leading dims / column offsets");
    SMTL_CHECK_ARG(a->head_dim == 0 || a->head_dim == 64 || (a->head_dim == 512 && a->heads == 1),
                   "fattn_plan: head_dim %d with %d heads (64 x any, or 512 x 1)", a->head_dim, a->heads);
    memset(op, 0, sizeof(*op));
    op->args = *a;
    if (a->head_dim == 512) {                       // the VAE mid-block attention: smtl_vattn_kernel
        op->grid_x = ((a->ntok + BQ - 1) / BQ + 1) / 2 * 2;       // CTA pairs (clusters of 2 along x)
        op->grid_y = a->batch;
        op->smem_bytes = V5_SMEM;
        const int rc = smtl_host::encode_tmap_bf16_2d(op->tmap_kv, a->qkv, (uint64_t)a->batch * a->ntok, (uint64_t)a->ld,
                                                      (uint64_t)a->ld, 64);
        if (rc != SMTL_OK) return rc;
    } else {
        op->grid_x = (a->ntok + F4_NQ * BQ - 1) / (F4_NQ * BQ);
        op->grid_y = a->batch * a->heads;
        op->smem_bytes = F4_SMEM;
    }
    SMTL_CHECK_ARG(op->grid_y <= 65535, "fattn_plan: batch*heads=%d exceeds grid.y", op->grid_y);
    return smtl_host::encode_tmap_bf16_2d(op->tmap_qkv, a->qkv, (uint64_t)a->batch * a->ntok, (uint64_t)a->ld,
                                          (uint64_t)a->ld, 128);
}

extern "C" int smtl_fattn_run(const smtl_fattn_op* op, void* stream) {
    SMTL_CHECK_ARG(op, "fattn_run: NULL op");
    const smtl_fattn_args& a = op->args;
    FattnKParams kp;
    memcpy(&kp.tm, op->tmap_qkv, 128);
    memcpy(&kp.tm_half, op->tmap_qkv, 128);
    kp.q_col0 = a.q_col0;
    kp.k_col0 = a.k_col0;
    kp.v_col0 = a.v_col0;
    kp.ntok = a.ntok;
    kp.heads = a.heads;
    kp.out = reinterpret_cast<uint16_t*>(a.out_bf16);
    kp.fmt = a.fmt16;
    kp.ldo = a.ldo;
    kp.scale_log2 = a.scale * 1.4426950408889634f;
    if (a.head_dim == 512) {
        static std::atomic<uint64_t> attr5{0};
        if (smtl_host::first_use_on_device(attr5))
            SMTL_CHECK_CUDA(cudaFuncSetAttribute(smtl_vattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, V5_SMEM));
        memcpy(&kp.tm_half, op->tmap_kv, 128);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(op->grid_x, op->grid_y, 1);
        cfg.blockDim = dim3(V5_THREADS, 1, 1);
        cfg.dynamicSmemBytes = op->smem_bytes;
        cfg.stream = reinterpret_cast<cudaStream_t>(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SMTL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, smtl_vattn_kernel, kp));
        SMTL_CHECK_CUDA(cudaGetLastError());
        return SMTL_OK;
    }
    static std::atomic<uint64_t> attr4{0};
    if (smtl_host::first_use_on_device(attr4)) {
        SMTL_CHECK_CUDA(cudaFuncSetAttribute(smtl_fattn4_kernel<FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM));
        SMTL_CHECK_CUDA(cudaFuncSetAttribute(smtl_fattn4_kernel<FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, F4_SMEM));
    }
    auto kern = a.fmt16 == SMTL_FMT_F16 ? smtl_fattn4_kernel<FMT_F16> : smtl_fattn4_kernel<FMT_BF16>;
    kern<<<dim3(op->grid_x, op->grid_y), F4_THREADS, op->smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(kp);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}
