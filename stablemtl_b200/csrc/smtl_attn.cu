// smtl_attn.cu -- flash-style self-attention for the SD-2 UNet transformer blocks (head dim 64).
// Replaces xformers.ops.memory_efficient_attention at reference src/model/attention.py:391-397,417.
//
// One CTA = 128 query rows of one (image, head); two CTAs are co-resident per SM so one CTA's softmax
// overlaps the other's tensor-core work.
//   warp 0     : TMA producer (Q once, then K/V tiles through a 2-stage ring)
//   warp 1     : MMA issuer   S = Q K^T (tcgen05, K-major operands)  and  O += P V (V read MN-major)
//   warps 2..5 : softmax      one query row per thread: tcgen05.ld S -> online softmax (fp32, exp2) ->
//                             P (bf16) into swizzled smem; rescale of O in TMEM; final O / l epilogue
#include "smtl_common.cuh"
#include "smtl_host.h"

namespace {
using namespace smtl;

constexpr int BQ = 128, BKV = 128, HD = 64;
constexpr int TILE_BYTES = 128 * 64 * 2;       // 16 KB: a [128 x 64] bf16 tile
constexpr int FA_THREADS = 192;
constexpr int FA_TMEM_COLS = 256;              // S: [0,128)  O: [128,192)
constexpr int FA_SMEM = 7 * TILE_BYTES + 256;  // Q, K0, K1, V0, V1, P0, P1 + barriers

struct alignas(64) FattnKParams {
    CUtensorMap tm;
    int32_t q_col0, k_col0, v_col0;
    int32_t ntok, heads;
    uint16_t* out;
    int32_t ldo;
    float scale_log2;
    int32_t fmt;
};

__global__ void __launch_bounds__(FA_THREADS, 2) smtl_fattn_kernel(const __grid_constant__ FattnKParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = smem + TILE_BYTES;          // 2 stages
    uint8_t* sV = smem + 3 * TILE_BYTES;      // 2 stages
    uint8_t* sP = smem + 5 * TILE_BYTES;      // 2 K-chunks of 64 kv columns
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * TILE_BYTES);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;     // [2]
    uint64_t* v_full = bars + 3;     // [2]
    uint64_t* kv_empty = bars + 5;   // [2]
    uint64_t* s_full = bars + 7;
    uint64_t* p_full = bars + 8;
    uint64_t* o_full = bars + 9;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * BQ;
    const int b = blockIdx.y / p.heads;
    const int hd = blockIdx.y - b * p.heads;
    const int row_base = b * p.ntok;
    const int ntiles = (p.ntok + BKV - 1) / BKV;

    if (threadIdx.x == 0) {
        if ((smem_u32(smem) & 1023u) != 0) { printf("smtl_fattn: smem base not 1024-aligned\n"); __trap(); }
        tma_prefetch_desc(&p.tm);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&k_full[s], 1); mbar_init(&v_full[s], 1); mbar_init(&kv_empty[s], 1); }
        mbar_init(s_full, 1);
        mbar_init(p_full, 4);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, FA_TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_s = tmem_base;
    const uint32_t tmem_o = tmem_base + 128;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, TILE_BYTES);
            tma_load_2d(sQ, &p.tm, q_full, p.q_col0 + hd * HD, row_base + q0);
            for (int j = 0; j < ntiles; ++j) {
                const int s = j & 1;
                mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1u);
                mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
                tma_load_2d(sK + s * TILE_BYTES, &p.tm, &k_full[s], p.k_col0 + hd * HD, row_base + j * BKV);
                mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
                tma_load_2d(sV + s * TILE_BYTES, &p.tm, &v_full[s], p.v_col0 + hd * HD, row_base + j * BKV);
            }
        }
    } else if (warp == 1) {
        const uint32_t IDESC_S = make_idesc_16(BQ, BKV, 0, 0, p.fmt);   // S[128,128] = Q[128,64] K[128,64]^T
        const uint32_t IDESC_O = make_idesc_16(BQ, HD, 0, 1, p.fmt);    // O[128,64] += P[128,128] V[128,64] (V MN-major)
        mbar_wait(q_full, 0);
        for (int j = 0; j < ntiles; ++j) {
            const int s = j & 1;
            const uint32_t kv_phase = (j >> 1) & 1;
            mbar_wait(&k_full[s], kv_phase);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t dq = make_smem_desc_sw128(smem_u32(sQ));
                const uint64_t dk = make_smem_desc_sw128(smem_u32(sK + s * TILE_BYTES));
#pragma unroll
                for (int k = 0; k < HD / 16; ++k) tc_mma_f16(tmem_s, dq + 2 * k, dk + 2 * k, IDESC_S, k != 0);
                tc_commit(s_full);
            }
            __syncwarp();
            mbar_wait(p_full, j & 1);      // softmax consumed S, wrote P, rescaled O
            mbar_wait(&v_full[s], kv_phase);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int kc = 0; kc < 2; ++kc) {
                    const uint64_t dp = make_smem_desc_sw128(smem_u32(sP + kc * TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        // V rows (kv) kc*64 + k*16 .. +16: 16 rows * 128 B = 2048 B per step
                        const uint64_t dv =
                            make_smem_desc_sw128(smem_u32(sV + s * TILE_BYTES) + (kc * 64 + k * 16) * 128);
                        tc_mma_f16(tmem_o, dp + 2 * k, dv, IDESC_O, (j | kc | k) != 0);
                    }
                }
                tc_commit(&kv_empty[s]);
                tc_commit(o_full);
            }
            __syncwarp();
        }
    } else {
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;                 // query row within the tile == TMEM lane
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f;
        uint8_t* prow = sP + (r >> 3) * 1024 + (r & 7) * 128;
        for (int j = 0; j < ntiles; ++j) {
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            const int kv_valid = p.ntok - j * BKV;         // columns >= kv_valid are out of range
            uint32_t rr[32];
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tmem_ld_32x32(tmem_s + lane_off + c * 32, rr);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float sv = (c * 32 + i < kv_valid) ? __uint_as_float(rr[i]) : -INFINITY;
                    mx = fmaxf(mx, sv);
                }
            }
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float alpha = exp2f(m_run - m_new);      // 0 on the first tile (m_run = -inf)
            // pass 2: p = exp2(s*scale - m_new) -> bf16 -> swizzled smem (A operand of the PV MMA)
            float psum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tmem_ld_32x32(tmem_s + lane_off + c * 32, rr);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    float p0 = (c * 32 + i < kv_valid) ? exp2f(__uint_as_float(rr[i]) * p.scale_log2 - m_new) : 0.f;
                    float p1 = (c * 32 + i + 1 < kv_valid) ? exp2f(__uint_as_float(rr[i + 1]) * p.scale_log2 - m_new) : 0.f;
                    const uint32_t u = pack16x2(p0, p1, p.fmt);
                    const float2 back = unpack16x2(u, p.fmt);   // sum what the tensor core will actually see
                    psum += back.x + back.y;
                    pk[i >> 1] = u;
                }
                uint8_t* chunk = prow + (c >> 1) * TILE_BYTES;
#pragma unroll
                for (int v4 = 0; v4 < 4; ++v4) {
                    const int c16 = (c & 1) * 4 + v4;      // 16-byte column within the 128-byte row
                    *reinterpret_cast<uint4*>(chunk + ((c16 ^ (r & 7)) << 4)) =
                        make_uint4(pk[4 * v4], pk[4 * v4 + 1], pk[4 * v4 + 2], pk[4 * v4 + 3]);
                }
            }
            l_run = l_run * alpha + psum;
            m_run = m_new;
            // rescale the running output (skipped when no row of this warp changed its max)
            if (j > 0) {
                const bool need = __any_sync(0xffffffffu, alpha != 1.0f);
                if (need) {
                    mbar_wait(o_full, (j - 1) & 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int c = 0; c < 2; ++c) {
                        tmem_ld_32x32(tmem_o + lane_off + c * 32, rr);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) rr[i] = __float_as_uint(__uint_as_float(rr[i]) * alpha);
                        tmem_st_32x32(tmem_o + lane_off + c * 32, rr);
                    }
                    tmem_st_wait();
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // epilogue: O / l -> bf16
        mbar_wait(o_full, (ntiles - 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l_run;
        const bool row_ok = (q0 + r) < p.ntok;
        uint16_t* dst = p.out + (int64_t)(row_base + q0 + r) * p.ldo + hd * HD;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t rr[32];
            tmem_ld_32x32(tmem_o + lane_off + c * 32, rr);
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
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, FA_TMEM_COLS);
    }
}

}  // namespace

extern "C" int smtl_fattn_plan(const smtl_fattn_args* a, smtl_fattn_op* op) {
    SMTL_CHECK_ARG(a && op && a->qkv && a->out_bf16, "fattn_plan: NULL argument");
    SMTL_CHECK_ARG(a->batch > 0 && a->ntok > 0 && a->heads > 0, "fattn_plan: empty problem");
    SMTL_CHECK_ARG(a->ld % 8 == 0 && a->ldo % 8 == 0 && a->q_col0 % 8 == 0 && a->k_col0 % 8 == 0 && a->v_col0 % 8 == 0,
                   "fattn_plan: unaligned leading dims / column offsets");
    memset(op, 0, sizeof(*op));
    op->args = *a;
    op->grid_x = (a->ntok + BQ - 1) / BQ;
    op->grid_y = a->batch * a->heads;
    SMTL_CHECK_ARG(op->grid_y <= 65535, "fattn_plan: batch*heads=%d exceeds grid.y", op->grid_y);
    op->smem_bytes = FA_SMEM;
    return smtl_host::encode_tmap_bf16_2d(op->tmap_qkv, a->qkv, (uint64_t)a->batch * a->ntok, (uint64_t)a->ld,
                                          (uint64_t)a->ld, 128);
}

extern "C" int smtl_fattn_run(const smtl_fattn_op* op, void* stream) {
    SMTL_CHECK_ARG(op, "fattn_run: NULL op");
    static bool attr_set = false;
    if (!attr_set) {
        SMTL_CHECK_CUDA(
            cudaFuncSetAttribute(smtl_fattn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM));
        attr_set = true;
    }
    const smtl_fattn_args& a = op->args;
    FattnKParams kp;
    memcpy(&kp.tm, op->tmap_qkv, 128);
    kp.q_col0 = a.q_col0;
    kp.k_col0 = a.k_col0;
    kp.v_col0 = a.v_col0;
    kp.ntok = a.ntok;
    kp.heads = a.heads;
    kp.out = reinterpret_cast<uint16_t*>(a.out_bf16);
    kp.fmt = a.fmt16;
    kp.ldo = a.ldo;
    kp.scale_log2 = a.scale * 1.4426950408889634f;
    smtl_fattn_kernel<<<dim3(op->grid_x, op->grid_y), FA_THREADS, op->smem_bytes, reinterpret_cast<cudaStream_t>(stream)>>>(kp);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}
