// smtl_gemm.cu -- the tensor-core workhorse of the StableMTL hot path.
//
// One persistent, warp-specialised tcgen05 kernel:
//   warp 0      : TMA producer  (cp.async.bulk.tensor 2-D, 128-byte swizzle, mbarrier complete_tx)
//   warp 1      : MMA issuer    (one elected lane issues tcgen05.mma, fp32 accumulators in TMEM, 2 stages)
//   warps 2..5  : epilogue      (tcgen05.ld -> bias / GELU / GEGLU / residual -> fp32 and/or bf16 stores)
//
// D[m, n] = sum over K segments of A_src[m + row_shift, a_col0 + k] * B[n, kbase + k].
// With one segment this is a plain token GEMM (nn.Linear, reference src/model/attention.py:185-212,442-460);
// with nine row-shifted segments over the zero-halo padded layout it is a 3x3 stride-1 convolution as an
// implicit GEMM (InflatedConv3d, src/model/resnet.py:14-16); a tenth segment from a second tensor fuses the
// 1x1 shortcut conv (src/model/resnet.py:172,200).
#include "smtl_common.cuh"
#include "smtl_host.h"

// The kernels are compiled in TWO translation units from this one file (build.sh / __graft_entry__.build()): the
// default one holds the host code and the fp16 instantiations, -DSMTL_GEMM_BF16_PART the bf16 instantiations behind
// this function -- 28 instantiations of the GEMM kernel in one ptxas run were three minutes of a four-file build.
// cg: 1 / 2 = smtl_gemm_kernel<block_n, cg, BF16>, 3 = smtl_gemmT_kernel<BF16>; kp: GemmKParams of the launch.
int smtl_detail_gemm_launch_bf16(int block_n, int cg, const void* kp, int grid, int smem_bytes, cudaStream_t stream);

namespace {

using namespace smtl;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;                       // one 128-byte swizzle atom of bf16
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 320;                   // TMA warp + MMA warp + 8 epilogue warps
constexpr int EPI_WARPS = 8;
constexpr int SMEM_BUDGET = 227 * 1024;
constexpr int GEMM_MAX_STAGES = 24;
// Statistics producers keep their per-column sums in shared memory across the tiles of one (image, column tile) run:
// [epilogue warp][32-column chunk of the warp][lane = column] x (sum, sum of squares) as int64 fixed point (fine cells).
constexpr int STATS_NCH = 4;                                       // chunks per warp: BN <= 256, two warps per lane quarter
constexpr int STATS_SMEM = EPI_WARPS * STATS_NCH * 32 * 2 * 8;     // 16 KB
// Shift-grouped mainloop: k-blocks per ring stage.  One barrier round trip (wait / expect_tx / TMA issue on one side, wait /
// MMAs / commit on the other) costs ~550 clk whatever the tile; with two 64-wide k-blocks per stage it is paid per 128 K.
constexpr int KPB = 2;
// The operand tiles are written by TMA and read by tcgen05.mma, both in the async proxy and ordered by the mbarrier's
// complete_tx; no tcgen05.fence is needed between the full-barrier wait and the MMAs (the fence after the acc_empty
// wait, which orders the epilogue's tcgen05.ld before the overwriting MMA, stays).
#ifdef SMTL_MAINLOOP_FENCE
#define MAINLOOP_FENCE() tc_fence_after()
#else
#define MAINLOOP_FENCE() ((void)0)
#endif                // mbarrier slots of smtl_gemm_kernel's rings (activation + weight)

struct alignas(64) GemmKParams {
    CUtensorMap tm_a0;
    CUtensorMap tm_a1;
    CUtensorMap tm_b;
    CUtensorMap tm_x8[2];  // swapped + grouped: 8-row boxes of a0 / a1 (a TMA box is at most 256 rows)
    int64_t m;
    int32_t n;            // B rows (pre-activation columns)
    int32_t n_out;        // output columns (n, or n/2 for GEGLU)
    int32_t tiles_m, tiles_n;
    int32_t total_kb;
    int32_t stages;
    int32_t nseg;
    smtl_gemm_seg seg[SMTL_MAX_SEG];
    const float* bias;
    int32_t bias_per_row;
    int32_t act;
    const void* res1;
    const void* res2;
    int32_t ldres;
    int32_t res16;
    unsigned long long* stats;   // int64 fixed-point cells [rep, images, n_out, 4] (smtl_common.cuh: stats_atomic_add)
    int32_t stats_rpi, stats_images, stats_rep, stats_g;
    // Image-aligned tiling (set whenever stats are produced): M tiles restart at every image's first GEMM row, so the
    // 32-row partial sums of an image are the same fp32 values wherever the image sits in the batch.
    int64_t tile_rpi;            // GEMM rows per image; 0 = tiles run over all rows
    int32_t tiles_per_img;
    // CTA pairs over image-aligned tiles of SMALL maps: the two CTAs of a pair take two CONSECUTIVE 128-row tiles of the
    // image-aligned 128-row tiling (tiles_per_img then counts 128-row tiles) instead of the halves of one 256-row tile
    // -- nothing in cta_group::2 needs the two halves of A to be adjacent rows.  A 15x20 map (374 padded rows) is 3 x 128
    // rows either way instead of 2 x 256 = 37 % padding, and keeps the pair's halved weight traffic (+17 % on these convs).
    int32_t pair_split;
    // Tile order.  0: tile = blockIdx + k * grid, column tile fastest.  1 (statistics producers): every CTA takes a
    // CONTIGUOUS run of the column-tile-major order, i.e. a long strip of M tiles of one column tile, so the per-column
    // sums stay in shared memory for a whole (image, column tile) run and reach global memory as one atomic per cell,
    // not one per 32-row slice.  CTAs c and c + grid / tiles_n walk the same rows at the same time (L2 reuse of A).
    int32_t tile_order;
    float* out_f32;
    uint16_t* out_bf16;
    uint16_t* aux_bf16;
    int32_t ldc;
    int32_t ld_aux;
    int32_t rowmap;
    int32_t img_h, img_w;
    int32_t up_py, up_px;
    int32_t fmt;
    int32_t lean;         // the launch qualifies for epilogue_rows_lean
    int64_t group_rows;   // 0: plain; else rows per weight group
    // "shift-grouped" mainloop: up to three K segments whose row shifts are consecutive (the kx = -1, 0, +1 taps of
    // one ky of a 3x3 conv) share ONE activation tile loaded with 8 extra rows; the MMA descriptor of tap j simply
    // starts j rows (j * 128 B) into it -- the 128-byte swizzle is a function of the shared-memory address, so a
    // row-shifted start reads the shifted rows (verified on B200).  A third of the activation TMA/L2/smem traffic.
    int32_t grouped;      // 0: one (A, B) ring; 1: activation ring + weight ring
    int32_t ngrp;
    int32_t sp, sw;       // stages of the activation / weight rings
    struct Grp { int32_t row_shift, nsub, kblocks, src, a_col0, kb0; } grp[SMTL_MAX_SEG];
    FastDiv fd_img, fd_row;   // row maps: rows per image (h*w or (h+2)*(w+2)) and per image row (w or w+2)
};

__device__ __forceinline__ void st_global_v4_f32(float* p, float a, float b, float c, float d) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void st_global_v4_b32(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// 256-bit global accesses (sm_100: STG.E.ENL2.256 / LDG.E.ENL2.256): a lane writes a whole 32-byte sector per instruction.
// With 16-byte row pieces every sector of the output was touched by two store instructions -- the SM -> L2 request
// rate, not DRAM, bound the short-K token linears.  Addresses must be 32-byte aligned.
__device__ __forceinline__ void st_global_v8_f32(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                 "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void st_global_v8_b32(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ldg_nc_v8(const void* p, uint32_t* v) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}
// 32 fp32 values -> 16-bit row piece (64 bytes), 256-bit stores when the row pitch allows
__device__ __forceinline__ void store_row16(uint16_t* dst, const float (&v)[32], int ld, int fmt) {   // fmt: compile-time at every call
    uint32_t w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = pack16x2(v[2 * j], v[2 * j + 1], fmt);
    if ((ld & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
        st_global_v8_b32(dst, w);
        st_global_v8_b32(dst + 16, w + 8);
    } else {
#pragma unroll
        for (int j = 0; j < 16; j += 4) st_global_v4_b32(dst + 2 * j, w[j], w[j + 1], w[j + 2], w[j + 3]);
    }
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// GEMM row -> output row for every SMTL_ROWMAP_* (see the header).  `halo` = a padded GEMM row that is not an interior
// pixel (PAD_KEEP stores zeros there; everyone else stores nothing).
// (rows fit in 31 bits -- checked by the plan -- and the two divisors are per-launch constants: multiply-shift division
// instead of 64-bit integer division, which was half of the swapped kernel's epilogue instructions)
__device__ __forceinline__ void map_row(const GemmKParams& p, int64_t grow, int64_t row_end, bool& ok, int64_t& orow, bool& halo, int& img_out) {
    ok = grow < row_end;
    orow = grow;
    halo = false;
    img_out = -1;                                            // identity map: the caller divides by stats_rpi itself
    if (p.rowmap == SMTL_ROWMAP_IDENTITY) return;
    const uint32_t g32 = (uint32_t)grow;
    if (p.rowmap == SMTL_ROWMAP_TO_PAD) {                    // compact pixel -> padded index
        const uint32_t hw = (uint32_t)(p.img_h * p.img_w);
        const int64_t img = fastdiv(g32, p.fd_img);
        const uint32_t rem = g32 - (uint32_t)img * hw;
        const int y = (int)fastdiv(rem, p.fd_row), x = (int)rem - y * p.img_w;
        orow = (img * (p.img_h + 2) + y + 1) * (int64_t)(p.img_w + 2) + x + 1;
        img_out = (int)img;
        return;
    }
    const int wp = p.img_w + 2;
    const uint32_t plane = (uint32_t)((p.img_h + 2) * wp);
    const int64_t img = fastdiv(g32, p.fd_img);
    const uint32_t rem = g32 - (uint32_t)img * plane;
    const int yp = (int)fastdiv(rem, p.fd_row), xp = (int)rem - yp * wp;
    img_out = (int)img;
    const bool interior = yp >= 1 && yp <= p.img_h && xp >= 1 && xp <= p.img_w;
    halo = ok && !interior;
    ok = ok && interior;
    if (p.rowmap == SMTL_ROWMAP_CONV_PAD)
        orow = (img * p.img_h + (yp - 1)) * p.img_w + (xp - 1);
    else if (p.rowmap == SMTL_ROWMAP_CONV_PAD_UP2)           // output parity (py, px) of the compact 2x map
        orow = (img * (2 * p.img_h) + 2 * (yp - 1) + p.up_py) * (int64_t)(2 * p.img_w) + 2 * (xp - 1) + p.up_px;
    else if (p.rowmap == SMTL_ROWMAP_UP2_PAD)                // ... of the padded 2x map
        orow = (img * (2 * p.img_h + 2) + 2 * (yp - 1) + p.up_py + 1) * (int64_t)(2 * p.img_w + 2) + 2 * (xp - 1) +
               p.up_px + 1;
    // PAD_KEEP: orow = grow
}

// First GEMM row of M tile `tm` (tile_rows = 128, 256 for a CTA pair or the swapped kernel) and the end of its valid rows.
__device__ __forceinline__ void tile_span(const GemmKParams& p, int tm, int tile_rows, int64_t& row0, int64_t& row_end) {
    if (p.tile_rpi) {
        const int img = tm / p.tiles_per_img;
        row0 = (int64_t)img * p.tile_rpi + (int64_t)(tm - img * p.tiles_per_img) * tile_rows;
        row_end = (int64_t)(img + 1) * p.tile_rpi;
        if (row_end > p.m) row_end = p.m;                  // the odd tile out of a split pair: no valid row
    } else {
        row0 = (int64_t)tm * tile_rows;
        row_end = p.m;
    }
}

// First GEMM row of the 128 rows CTA `rank` of a pair (or the only CTA) works on in M tile `tm`, and the end of its valid rows.
template <int CG>
__device__ __forceinline__ void cta_span(const GemmKParams& p, int tm, uint32_t rank, int64_t& row0, int64_t& row_end) {
    if (CG == 2 && p.pair_split) {
        tile_span(p, 2 * tm + (int)rank, BLOCK_M, row0, row_end);
    } else {
        tile_span(p, tm, CG * BLOCK_M, row0, row_end);
        row0 += (int64_t)rank * BLOCK_M;
    }
}

// Per-thread state of one tile's epilogue, computed BEFORE waiting for the accumulator so that the residual
// prefetches and the bias loads overlap the tile's MMAs.
template <int BN>
struct EpiRow {
    int64_t orow;          // output row of this thread's accumulator row
    bool row_ok;           // false: halo / out-of-range row (nothing stored)
    bool halo;             // PAD_KEEP: this halo row of the output is zeroed
    int img;               // image of the output row (statistics), -1 if !row_ok
    int img_lo, img_hi;    // warp-wide range of img over valid rows (img_lo > img_hi: no valid row)
    float row_bias;
    int64_t bias_ofs;      // grouped GEMM: first bias element of this row's weight group (0 otherwise)
    float bias[BN / 32];   // lane-distributed: bias[k] = bias_vec[n0 + 32 k + lane]
};

template <int BN>
__device__ __forceinline__ void epilogue_prepare(const GemmKParams& p, int64_t grow, int64_t row_end, int tn, int lane, EpiRow<BN>& e) {
    const bool geglu = (p.act == SMTL_ACT_GEGLU);
    const int out_bn = geglu ? BN / 2 : BN;
    const int n0 = tn * BN;
    int map_img;
    map_row(p, grow, row_end, e.row_ok, e.orow, e.halo, map_img);
    e.halo = e.halo && (p.rowmap == SMTL_ROWMAP_PAD_KEEP);
    e.row_bias = (p.bias && p.bias_per_row && grow < p.m) ? __ldg(p.bias + grow) : 0.0f;
    // this tile's weight group.  Rows past the end of the tile's valid rows (the tail of a group that is not a whole
    // number of tiles -- past the end of the MATRIX for the last group) take the group of the last valid row: their
    // results are never stored, but the bias they index must exist.
    const int64_t grow_c = grow < row_end ? grow : row_end - 1;
    e.bias_ofs = p.group_rows ? (grow_c / p.group_rows) * p.n : 0;
#pragma unroll
    for (int k = 0; k < BN / 32; ++k) {
        const int c = n0 + 32 * k + lane;
        const int64_t gofs = e.bias_ofs;
        e.bias[k] = (p.bias && !p.bias_per_row && c < p.n) ? __ldg(p.bias + gofs + c) : 0.0f;
    }
    // pull this row's residual lines into L2 while the MMAs of the tile run
    if (e.row_ok && (p.res1 || p.res2)) {
        const int ocol0 = geglu ? tn * (BN / 2) : n0;
        const int esz = p.res16 ? 2 : 4;
        int ncols = p.n_out - ocol0;
        if (ncols > out_bn) ncols = out_bn;
        const int64_t off = (e.orow * (int64_t)p.ldres + ocol0) * esz;
        for (int b = 0; b < ncols * esz; b += 128) {
            if (p.res1) prefetch_l2(reinterpret_cast<const char*>(p.res1) + off + b);
            if (p.res2) prefetch_l2(reinterpret_cast<const char*>(p.res2) + off + b);
        }
    }
    e.img = -1;
    e.img_lo = 1;
    e.img_hi = 0;
    if (p.stats) {
        e.img = !e.row_ok ? -1 : (map_img >= 0 ? map_img : (int)(e.orow / p.stats_rpi));
        int lo = e.row_ok ? e.img : 0x7fffffff, hi = e.img;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        e.img_lo = lo;
        e.img_hi = hi;
    }
}

// residual add of one 32-column chunk of this thread's row (fp32 or 16-bit source).
// FMT (and with it every conversion below) is a compile-time constant: with a run-time format ptxas emits BOTH
// conversion sequences under predicates, and a predicated-off instruction still takes its issue slot.
template <int FMT>
__device__ __forceinline__ void add_residual(const GemmKParams& p, const void* res, int64_t orow, int ocol, bool full,
                                             float (&v)[32]) {
    if (p.res16) {
        const uint16_t* src = reinterpret_cast<const uint16_t*>(res) + orow * (int64_t)p.ldres + ocol;
        if (full && (p.ldres & 15) == 0 && (reinterpret_cast<uintptr_t>(src) & 31) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 16) {
                uint32_t t[8];
                ldg_nc_v8(src + j, t);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float2 f = unpack16x2(t[q], FMT);
                    v[j + 2 * q] += f.x;
                    v[j + 2 * q + 1] += f.y;
                }
            }
        } else if (full && (p.ldres & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                const uint4 t = ldg_nc_v4(src + j);
                float2 f;
                f = unpack16x2(t.x, FMT); v[j] += f.x; v[j + 1] += f.y;
                f = unpack16x2(t.y, FMT); v[j + 2] += f.x; v[j + 3] += f.y;
                f = unpack16x2(t.z, FMT); v[j + 4] += f.x; v[j + 5] += f.y;
                f = unpack16x2(t.w, FMT); v[j + 6] += f.x; v[j + 7] += f.y;
            }
        } else {
            for (int j = 0; j < 32; ++j)
                if (ocol + j < p.n_out) v[j] += unpack16x2((uint32_t)src[j], FMT).x;
        }
    } else {
        const float* src = reinterpret_cast<const float*>(res) + orow * (int64_t)p.ldres + ocol;
        if (full && (p.ldres & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 31) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint32_t t[8];
                ldg_nc_v8(src + j, t);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[j + q] += __uint_as_float(t[q]);
            }
        } else if (full && (p.ldres & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(src + j));
                v[j] += t.x; v[j + 1] += t.y; v[j + 2] += t.z; v[j + 3] += t.w;
            }
        } else {
            for (int j = 0; j < 32; ++j)
                if (ocol + j < p.n_out) v[j] += __ldg(src + j);
        }
    }
}

// The statistics of one chunk (lane = row, v = the 32 stored columns starting at `ocol`) into this warp's on-chip cells.
__device__ __forceinline__ void chunk_stats_to_cells(const GemmKParams& p, const float (&v)[32], bool row_ok, int lane, int ocol,
                                                     int c0, int img, uint32_t stats_acc) {
    // stats_g adjacent channels share a cell when every consumer's GroupNorm groups are unions of such blocks
    // (smtl_gemm_args.stats_group): the cross-lane reduction shrinks from 2 x 31 shuffles to 2 x 6 at 8 channels
    float cs, cq;
    if (p.stats_g == 8) chunk_col_stats<8>(v, row_ok, lane, cs, cq);
    else if (p.stats_g == 4) chunk_col_stats<4>(v, row_ok, lane, cs, cq);
    else if (p.stats_g == 2) chunk_col_stats<2>(v, row_ok, lane, cs, cq);
    else chunk_col_stats<1>(v, row_ok, lane, cs, cq);
    if (ocol + lane < p.n_out && (lane & (p.stats_g - 1)) == 0) {
        // fine cells accumulate in this warp's shared-memory slots (flushed once per (image, column tile) run);
        // a partial too large for the fine scale (|x| >= 2^14: rare) goes straight to its coarse global cell
        const uint32_t cell = stats_acc + (uint32_t)(((c0 >> 6) * 32 + lane) * 16);
        unsigned long long* gcell = p.stats + (((int64_t)(blockIdx.x % p.stats_rep) * p.stats_images + img) *
                                                   p.n_out + ocol + lane) * 4;
        long long as, aq;
        lds_v2_s64(cell, as, aq);
        bool hi;
        long long f = stats_fix(cs, hi);
        if (hi) atomicAdd(gcell + 1, (unsigned long long)f); else as += f;
        f = stats_fix(cq, hi);
        if (hi) atomicAdd(gcell + 3, (unsigned long long)f); else aq += f;
        sts_v2_s64(cell, as, aq);
    }
}

// The epilogue of the path's commonest GEMM, with nothing else in the instruction stream: column bias, optional 16-bit
// residual, 16-bit output (+ zero halo), GroupNorm statistics -- every 3x3 conv of the VAE and the UNet and the plain
// token linears.  The general epilogue below serves that case with ~830 instructions per 32 x 32 chunk (ncu: 47 branches
// on per-launch flags, 32 shuffles to broadcast the bias, scalar fallbacks in every store) and two warps per SM
// sub-partition to issue them: a K = 1024 conv spent 10 us in the epilogue of a tile whose MMAs take 6.4.
template <int BN, int FMT, int ACT>
__device__ __forceinline__ void epilogue_rows_lean(const GemmKParams& p, uint32_t taddr, int tn, int lane, int half,
                                                   const EpiRow<BN>& e, uint32_t stats_acc) {
    constexpr bool GEGLU = ACT == SMTL_ACT_GEGLU;
    constexpr int OUT_BN = GEGLU ? BN / 2 : BN;            // GEGLU: [value 128 | gate 128] per tile -> 128 output columns
    const int n0 = tn * BN;
    const bool row_ok = e.row_ok;
    const float* const bias = p.bias ? p.bias + e.bias_ofs : nullptr;
    const bool do_stats = ACT == SMTL_ACT_NONE && p.stats && e.img_lo <= e.img_hi;
#pragma unroll 1
    for (int c0 = half * 32; c0 < OUT_BN; c0 += 64) {
        const int ncol_in = n0 + c0;                       // B-row index of the chunk's first column
        if (ncol_in >= p.n) break;                         // warp-uniform
        const int ocol = GEGLU ? tn * OUT_BN + c0 : ncol_in;
        const bool whole = GEGLU || ocol + 32 <= p.n_out;  // else the last chunk of a launch with n % 32 == 16: 16 columns
        uint32_t r[32];
        float v[32];
        tmem_ld_32x32(taddr + c0, r);
        // the chunk's 32 bias values: every lane reads the SAME 16-byte pieces (one broadcast transaction each, L1-resident)
        float4 b4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
            b4[i] = (bias && (whole || i < 4)) ? __ldg(reinterpret_cast<const float4*>(bias + ncol_in) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t rs[16];                                   // 16-bit residual, issued before the TMEM wait
        const bool res16 = ACT == SMTL_ACT_NONE && p.res1 && p.res16 && row_ok;
        if (res16) {
            const uint16_t* src = reinterpret_cast<const uint16_t*>(p.res1) + e.orow * (int64_t)p.ldres + ocol;
            ldg_nc_v8(src, rs);
            if (whole) {
                ldg_nc_v8(src + 16, rs + 8);
            } else {
#pragma unroll
                for (int q = 8; q < 16; ++q) rs[q] = 0u;
            }
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            v[4 * i] = __uint_as_float(r[4 * i]) + b4[i].x;
            v[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b4[i].y;
            v[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b4[i].z;
            v[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b4[i].w;
        }
        if (GEGLU) {                                       // v = value * gelu(gate)
            tmem_ld_32x32(taddr + OUT_BN + c0, r);
            if (bias) {
#pragma unroll
                for (int i = 0; i < 8; ++i) b4[i] = __ldg(reinterpret_cast<const float4*>(bias + ncol_in + OUT_BN) + i);
            }
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 g0 = gelu_poly2(make_float2(__uint_as_float(r[4 * i]) + b4[i].x, __uint_as_float(r[4 * i + 1]) + b4[i].y));
                const float2 g1 = gelu_poly2(make_float2(__uint_as_float(r[4 * i + 2]) + b4[i].z, __uint_as_float(r[4 * i + 3]) + b4[i].w));
                v[4 * i] *= g0.x; v[4 * i + 1] *= g0.y; v[4 * i + 2] *= g1.x; v[4 * i + 3] *= g1.y;
            }
        } else if (ACT == SMTL_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 g = gelu_poly2(make_float2(v[j], v[j + 1]));
                v[j] = g.x;
                v[j + 1] = g.y;
            }
        }
        if (row_ok) {
            if (p.aux_bf16) {                              // the value before the residual add (child-stream feature tap)
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = pack16x2(v[2 * j], v[2 * j + 1], FMT);
                uint16_t* dst = p.aux_bf16 + e.orow * (int64_t)p.ld_aux + ocol;
                st_global_v8_b32(dst, w);
                if (whole) st_global_v8_b32(dst + 16, w + 8);
            }
            if (res16) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float2 f = unpack16x2(rs[q], FMT);
                    v[2 * q] += f.x;
                    v[2 * q + 1] += f.y;
                }
            } else if (p.res1) {                           // fp32 residual stream(s) of the transformer
#pragma unroll 1
                for (int k = 0; k < 2; ++k) {
                    const float* res = reinterpret_cast<const float*>(k == 0 ? p.res1 : p.res2);
                    if (!res) break;
                    const float* src = res + e.orow * (int64_t)p.ldres + ocol;
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        if (!whole && j >= 16) break;
                        uint32_t t[8];
                        ldg_nc_v8(src + j, t);
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[j + q] += __uint_as_float(t[q]);
                    }
                }
            }
            if (p.out_f32) {
                float* dst = p.out_f32 + e.orow * (int64_t)p.ldc + ocol;
#pragma unroll
                for (int j = 0; j < 32; j += 8)
                    if (whole || j < 16) st_global_v8_f32(dst + j, &v[j]);
            }
            if (p.out_bf16) {
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = pack16x2(v[2 * j], v[2 * j + 1], FMT);
                uint16_t* dst = p.out_bf16 + e.orow * (int64_t)p.ldc + ocol;
                st_global_v8_b32(dst, w);
                if (whole) st_global_v8_b32(dst + 16, w + 8);
            }
        } else if (e.halo && p.out_bf16) {                 // PAD_KEEP: the output keeps a zero halo for its consumers
            const uint32_t z[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
            uint16_t* dst = p.out_bf16 + e.orow * (int64_t)p.ldc + ocol;
            st_global_v8_b32(dst, z);
            if (whole) st_global_v8_b32(dst + 16, z);
        }
        __syncwarp();   // reconverge before the next warp-collective instruction
        if (do_stats) chunk_stats_to_cells(p, v, row_ok, lane, ocol, c0, e.img_lo, stats_acc);
    }
}

// Epilogue for one accumulator row per thread: `taddr` = TMEM address of this warp's lane quarter, column 0 of the
// tile; `tn` = N-tile index.  All tcgen05.ld / shuffles are warp-collective.
template <int BN, int FMT>
__device__ __forceinline__ void epilogue_rows(const GemmKParams& p, uint32_t taddr, int tn, int lane, int half,
                                              const EpiRow<BN>& e, uint32_t stats_acc) {
    if (p.lean) {                                          // warp-uniform, fixed per launch (smtl_gemm_run)
        if (p.act == SMTL_ACT_GEGLU) epilogue_rows_lean<BN, FMT, SMTL_ACT_GEGLU>(p, taddr, tn, lane, half, e, stats_acc);
        else if (p.act == SMTL_ACT_GELU) epilogue_rows_lean<BN, FMT, SMTL_ACT_GELU>(p, taddr, tn, lane, half, e, stats_acc);
        else epilogue_rows_lean<BN, FMT, SMTL_ACT_NONE>(p, taddr, tn, lane, half, e, stats_acc);
        return;
    }
    const bool geglu = (p.act == SMTL_ACT_GEGLU);
    const int out_bn = geglu ? BN / 2 : BN;
    const int n0 = tn * BN;
    const bool row_ok = e.row_ok;
    const int64_t orow = e.orow;
    float bq[BN / 32];          // bias queue: bq[0] is always the current chunk's (rotated once per chunk, so the
#pragma unroll                  // array is only ever indexed with compile-time constants and stays in registers)
    for (int k = 0; k < BN / 32; ++k) bq[k] = e.bias[k];
    if (half) {                 // the two warps of a lane quarter take alternate 32-column chunks
#pragma unroll
        for (int k = 0; k + 1 < BN / 32; ++k) bq[k] = bq[k + 1];
    }

#pragma unroll 1
    for (int c0 = half * 32; c0 < out_bn; c0 += 64) {
        const int ncol_in = n0 + c0;                       // B-row index of the chunk's first column
        if (ncol_in >= p.n) break;                         // warp-uniform
        uint32_t r[32];
        float v[32];
        tmem_ld_32x32(taddr + c0, r);
        tmem_ld_wait();
        if (p.bias_per_row) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + e.row_bias;
        } else {
            const float bl = bq[0];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bl, j);
        }
        int ocol = ncol_in;                                // output column of v[0]
        if (geglu) {
            tmem_ld_32x32(taddr + BN / 2 + c0, r);
            tmem_ld_wait();
            const float gl = bq[BN / 64];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 g = gelu_poly2(make_float2(__uint_as_float(r[j]) + __shfl_sync(0xffffffffu, gl, j),
                                                        __uint_as_float(r[j + 1]) + __shfl_sync(0xffffffffu, gl, j + 1)));
                v[j] *= g.x;
                v[j + 1] *= g.y;
            }
            ocol = tn * (BN / 2) + c0;
        } else if (p.act == SMTL_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const float2 g = gelu_poly2(make_float2(v[j], v[j + 1]));
                v[j] = g.x;
                v[j + 1] = g.y;
            }
        } else if (p.act == SMTL_ACT_SILU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
        }
        const bool full = (ocol + 32 <= p.n_out);
        if (row_ok) {
            if (p.aux_bf16) {
                uint16_t* dst = p.aux_bf16 + orow * (int64_t)p.ld_aux + ocol;
                if (full && (p.ld_aux & 7) == 0) {
                    store_row16(dst, v, p.ld_aux, FMT);
                } else {
                    for (int j = 0; j < 32; ++j)
                        if (ocol + j < p.n_out) dst[j] = to16(v[j], FMT);
                }
            }
            if (p.res1) add_residual<FMT>(p, p.res1, orow, ocol, full, v);
            if (p.res2) add_residual<FMT>(p, p.res2, orow, ocol, full, v);
            if (p.out_f32) {
                float* dst = p.out_f32 + orow * (int64_t)p.ldc + ocol;
                if (full && (p.ldc & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) st_global_v8_f32(dst + j, &v[j]);
                } else if (full && (p.ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) st_global_v4_f32(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
                    for (int j = 0; j < 32; ++j)
                        if (ocol + j < p.n_out) dst[j] = v[j];
                }
            }
            if (p.out_bf16) {
                uint16_t* dst = p.out_bf16 + orow * (int64_t)p.ldc + ocol;
                if (full && (p.ldc & 7) == 0) {
                    store_row16(dst, v, p.ldc, FMT);
                } else {
                    for (int j = 0; j < 32; ++j)
                        if (ocol + j < p.n_out) dst[j] = to16(v[j], FMT);
                }
            }
        }
        if (e.halo && p.out_bf16) {                 // PAD_KEEP: the output keeps a zero halo for its consumers
            uint16_t* dst = p.out_bf16 + orow * (int64_t)p.ldc + ocol;
            if (full && (p.ldc & 7) == 0) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) st_global_v4_b32(dst + j, 0u, 0u, 0u, 0u);
            } else {
                for (int j = 0; j < 32; ++j)
                    if (ocol + j < p.n_out) dst[j] = 0;
            }
        }
        __syncwarp();   // reconverge before the next warp-collective instruction
#pragma unroll
        for (int k = 0; k + 2 < BN / 32; ++k) bq[k] = bq[k + 2];
        if (p.stats && e.img_lo <= e.img_hi) {
            // per-(image, channel) sum / sum of squares of the stored value: lane j ends up owning column ocol + j.
            // Tiles are image-aligned (tile_rpi), so every valid row of the tile belongs to image e.img_lo.
            chunk_stats_to_cells(p, v, row_ok, lane, ocol, c0, e.img_lo, stats_acc);
        }
    }
}

// CG = 1: one CTA per 128-row tile.  CG = 2: a CTA PAIR (2-SM cluster) per 256-row tile -- each CTA stages its own
// 128 A rows and HALF of the B tile, the leader issues tcgen05.mma.cta_group::2 (M = 256) and each CTA's TMEM
// receives its 128 accumulator rows: half the shared-memory and L2 operand traffic per flop.
template <int BN, int CG, int FMT>
__global__ void __launch_bounds__(NUM_THREADS, 1) smtl_gemm_kernel(const __grid_constant__ GemmKParams p) {
    constexpr int B_ROWS = BN / CG;                                   // B rows this CTA stages
    constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
    constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    constexpr int TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    constexpr int ACC_STRIDE = TMEM_COLS / 2;
    constexpr int MAX_STAGES = GEMM_MAX_STAGES;

    constexpr int PB = (BLOCK_M + 8) * BLOCK_K * 2;                  // grouped mode: activation tile with 8 extra rows
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    const bool grouped = p.grouped != 0;
    constexpr int PB2 = KPB * PB, WB2 = KPB * B_STAGE_BYTES;           // grouped mode: KPB k-blocks per ring stage
    uint8_t* smem_w = smem + (size_t)p.sp * PB2;                      // grouped mode: weight ring after the activation ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(
        grouped ? smem_w + (size_t)p.sw * WB2 : smem + (size_t)stages * STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [stages]   TMA (both CTAs) -> MMA (leader)
    uint64_t* empty_bar = bars + MAX_STAGES;         // [stages]   MMA -> TMA (each CTA its own)
    uint64_t* acc_full = bars + 2 * MAX_STAGES;      // [2]        MMA -> epilogue (each CTA its own)
    uint64_t* acc_empty = bars + 2 * MAX_STAGES + 2; // [2]        epilogue (both CTAs) -> MMA (leader)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = (rank == 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tm_a0);
        tma_prefetch_desc(&p.tm_a1);
        tma_prefetch_desc(&p.tm_b);
        for (int s = 0; s < MAX_STAGES; ++s) {     // grouped mode: [0, sp) activation ring, [sp, sp + sw) weight ring
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], EPI_WARPS * CG);   // one arrive per epilogue warp of every CTA of the pair
        }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);       // each CTA allocates its whole accumulator space: same address in both
        tmem_relinquish();
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int num_tiles = p.tiles_m * p.tiles_n;     // tiles_m counts (128 * CG)-row blocks
    int tile_begin = blockIdx.x / CG, tile_end = num_tiles, tile_inc = gridDim.x / CG;
    if (p.tile_order) {                              // contiguous run of the column-tile-major order
        const int nslots = gridDim.x / CG, slot = blockIdx.x / CG;
        tile_begin = (int)((int64_t)slot * num_tiles / nslots);
        tile_end = (int)((int64_t)(slot + 1) * num_tiles / nslots);
        tile_inc = 1;
    }
    auto decode_tile = [&](int tile, int& tm, int& tn) {
        if (p.tile_order) { tn = tile / p.tiles_m; tm = tile - tn * p.tiles_m; }
        else { tm = tile / p.tiles_n; tn = tile - tm * p.tiles_n; }
    };

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        // converged warp, elect.sync at the TMA issue (same reason as for the MMA warp below)
        if (grouped) {
            int ps = 0, ws = 0;
            uint32_t pph = 0, wph = 0;
            for (int tile = tile_begin; tile < tile_end; tile += tile_inc) {
                int tm, tn;
                decode_tile(tile, tm, tn);
                int64_t m0, row_end;
                cta_span<CG>(p, tm, rank, m0, row_end);
                const int n0 = tn * BN + (int)rank * B_ROWS +
                               (p.group_rows ? (int)(m0 / p.group_rows) * p.n : 0);       // grouped GEMMs are single-CTA
                for (int g = 0; g < p.ngrp; ++g) {
                    const GemmKParams::Grp gr = p.grp[g];
                    const CUtensorMap* tma = gr.src ? &p.tm_a1 : &p.tm_a0;
                    for (int kb = 0; kb < gr.kblocks; kb += KPB) {
                        const int nk = min(KPB, gr.kblocks - kb);          // k-blocks in this stage (the last may be short)
                        mbar_wait(&empty_bar[ps], pph ^ 1u);
                        if (elect_one()) {
                            uint8_t* sa = smem + (size_t)ps * PB2;
                            const int32_t arow = (int32_t)(m0 + gr.row_shift);
                            if (CG == 1) {
                                mbar_arrive_expect_tx(&full_bar[ps], nk * PB);
                                for (int j = 0; j < nk; ++j)
                                    tma_load_2d(sa + j * PB, tma, &full_bar[ps], gr.a_col0 + (kb + j) * BLOCK_K, arow);
                            } else {        // both CTAs' tiles complete on the LEADER's barrier
                                if (leader) mbar_arrive_expect_tx(&full_bar[ps], 2 * nk * PB);
                                const uint32_t bar = mapa_u32(&full_bar[ps], 0);
                                for (int j = 0; j < nk; ++j)
                                    tma_load_2d_pair(sa + j * PB, tma, bar, gr.a_col0 + (kb + j) * BLOCK_K, arow);
                            }
                        }
                        __syncwarp();
                        if (++ps == p.sp) { ps = 0; pph ^= 1u; }
                        for (int sub = 0; sub < gr.nsub; ++sub) {
                            const int wb = p.sp + ws;
                            mbar_wait(&empty_bar[wb], wph ^ 1u);
                            if (elect_one()) {
                                uint8_t* sb = smem_w + (size_t)ws * WB2;
                                const int kcol = (gr.kb0 + sub * gr.kblocks + kb) * BLOCK_K;
                                if (CG == 1) {
                                    mbar_arrive_expect_tx(&full_bar[wb], nk * B_STAGE_BYTES);
                                    for (int j = 0; j < nk; ++j)
                                        tma_load_2d(sb + j * B_STAGE_BYTES, &p.tm_b, &full_bar[wb], kcol + j * BLOCK_K, n0);
                                } else {
                                    if (leader) mbar_arrive_expect_tx(&full_bar[wb], 2 * nk * B_STAGE_BYTES);
                                    const uint32_t bar = mapa_u32(&full_bar[wb], 0);
                                    for (int j = 0; j < nk; ++j)
                                        tma_load_2d_pair(sb + j * B_STAGE_BYTES, &p.tm_b, bar, kcol + j * BLOCK_K, n0);
                                }
                            }
                            __syncwarp();
                            if (++ws == p.sw) { ws = 0; wph ^= 1u; }
                        }
                    }
                }
            }
        } else {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile_begin; tile < tile_end; tile += tile_inc) {
                int tm, tn;
                decode_tile(tile, tm, tn);
                int64_t m0, row_end;
                cta_span<CG>(p, tm, rank, m0, row_end);
                const int n0 = tn * BN + (int)rank * B_ROWS +
                               (p.group_rows ? (int)(m0 / p.group_rows) * p.n : 0);       // grouped GEMMs are single-CTA
                int kb_global = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const smtl_gemm_seg sg = p.seg[s];
                    const CUtensorMap* tma = sg.src ? &p.tm_a1 : &p.tm_a0;
                    const int32_t arow = (int32_t)(m0 + sg.row_shift);
                    for (int kb = 0; kb < sg.kblocks; ++kb, ++kb_global) {
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        uint8_t* sa = smem + (size_t)stage * STAGE_BYTES;
                        uint8_t* sb = sa + A_STAGE_BYTES;
                        if (elect_one()) {
                            if (CG == 1) {
                                mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                                tma_load_2d(sa, tma, &full_bar[stage], sg.a_col0 + kb * BLOCK_K, arow);
                                tma_load_2d(sb, &p.tm_b, &full_bar[stage], kb_global * BLOCK_K, n0);
                            } else {
                                // both CTAs' loads complete on the LEADER's barrier, which expects the bytes of both
                                if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
                                const uint32_t bar = mapa_u32(&full_bar[stage], 0);
                                tma_load_2d_pair(sa, tma, bar, sg.a_col0 + kb * BLOCK_K, arow);
                                tma_load_2d_pair(sb, &p.tm_b, bar, kb_global * BLOCK_K, n0);
                            }
                        }
                        __syncwarp();
                        if (++stage == stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        // The whole warp runs the loop CONVERGED and elect.sync picks the issuing lane right at the instructions (the
        // CUTLASS idiom).  Under a lane-id branch (`if (lane == 0)`) ptxas cannot prove that the operands of UTCHMMA /
        // UTCBAR are warp-uniform and wraps every MMA in an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop: ~165 clk
        // of issue overhead per MMA (ncu source view), a floor above the 16..128 tensor cycles of the MMA itself.
        if (leader && grouped) {
            const uint32_t IDESC = make_idesc_16(BLOCK_M * CG, BN, 0, 0, FMT);
            int ps = 0, ws = 0;
            uint32_t pph = 0, wph = 0;
            int it = 0;
            for (int tile = tile_begin; tile < tile_end; tile += tile_inc, ++it) {
                const int acc = it & 1;
                mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * ACC_STRIDE;
                uint32_t accum = 0;
                for (int g = 0; g < p.ngrp; ++g) {
                    const int nsub = p.grp[g].nsub, nkb = p.grp[g].kblocks;
                    for (int kb = 0; kb < nkb; kb += KPB) {
                        const int nk = min(KPB, nkb - kb);
                        mbar_wait(&full_bar[ps], pph);
                        const uint32_t sa = smem_u32(smem + (size_t)ps * PB2);
                        for (int sub = 0; sub < nsub; ++sub) {
                            const int wb = p.sp + ws;
                            mbar_wait(&full_bar[wb], wph);
                            MAINLOOP_FENCE();
                            const uint32_t sb = smem_u32(smem_w + (size_t)ws * WB2);
                            if (elect_one()) {
                                for (int j = 0; j < nk; ++j) {
                                    const uint64_t da = make_smem_desc_sw128(sa + j * PB + sub * 128);  // tap `sub`: one row further
                                    const uint64_t db = make_smem_desc_sw128(sb + j * B_STAGE_BYTES);
#pragma unroll
                                    for (int k = 0; k < BLOCK_K / 16; ++k) {
                                        if (CG == 1)
                                            tc_mma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, accum | (uint32_t)(j | k));
                                        else
                                            tc_mma_f16_pair(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, accum | (uint32_t)(j | k));
                                    }
                                }
                                if (CG == 1) tc_commit(&empty_bar[wb]); else tc_commit_pair(&empty_bar[wb]);
                            }
                            accum = 1;
                            __syncwarp();
                            if (++ws == p.sw) { ws = 0; wph ^= 1u; }
                        }
                        if (elect_one()) { if (CG == 1) tc_commit(&empty_bar[ps]); else tc_commit_pair(&empty_bar[ps]); }
                        __syncwarp();
                        if (++ps == p.sp) { ps = 0; pph ^= 1u; }
                    }
                }
                if (elect_one()) { if (CG == 1) tc_commit(&acc_full[acc]); else tc_commit_pair(&acc_full[acc]); }
                __syncwarp();
            }
        } else if (leader) {
            const uint32_t IDESC = make_idesc_16(BLOCK_M * CG, BN, 0, 0, FMT);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = tile_begin; tile < tile_end; tile += tile_inc, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&acc_empty[acc], acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * ACC_STRIDE;
                for (int kb = 0; kb < p.total_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    MAINLOOP_FENCE();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE_BYTES);
                    const uint64_t da = make_smem_desc_sw128(sa);
                    const uint64_t db = make_smem_desc_sw128(sa + A_STAGE_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; ++k) {
                            // +32 B per 16-element K step inside the swizzle atom: start-address field += 2
                            if (CG == 1)
                                tc_mma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) != 0);
                            else
                                tc_mma_f16_pair(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC,
                                                (kb | k) != 0);
                        }
                        if (CG == 1) {
                            tc_commit(&empty_bar[stage]);                       // smem slot free once these MMAs retire
                            if (kb == p.total_kb - 1) tc_commit(&acc_full[acc]); // accumulator ready
                        } else {
                            tc_commit_pair(&empty_bar[stage]);
                            if (kb == p.total_kb - 1) tc_commit_pair(&acc_full[acc]);
                        }
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9)
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;             // which of the quarter's two warps: alternate column chunks
        const int row_in_tile = quarter * 32 + lane;
        // this warp's running column sums (see STATS_SMEM) and the (image, column tile) run they belong to
        const uint32_t stats_acc = smem_u32(reinterpret_cast<uint8_t*>(bars) + 512) +
                                   (uint32_t)((warp - 2) * STATS_NCH * 32 * 16);        // [chunk][lane] x (sum, sq) int64
        int run_img = -1, run_tn = -1;
        if (p.stats) {
#pragma unroll
            for (int i = 0; i < STATS_NCH; ++i) sts_v2_s64(stats_acc + (uint32_t)((i * 32 + lane) * 16), 0, 0);
            __syncwarp();
        }
        auto stats_flush = [&](int img, int tn) {
            if (img < 0) return;
#pragma unroll
            for (int i = 0; i < STATS_NCH; ++i) {
                const int col = tn * BN + half * 32 + 64 * i + lane;
                const uint32_t cell = stats_acc + (uint32_t)((i * 32 + lane) * 16);
                long long fs, fq;
                lds_v2_s64(cell, fs, fq);
                if (half * 32 + 64 * i < BN && col < p.n_out && (fs | fq)) {
                    unsigned long long* g = p.stats + (((int64_t)(blockIdx.x % p.stats_rep) * p.stats_images + img) * p.n_out + col) * 4;
                    atomicAdd(g, (unsigned long long)fs);
                    atomicAdd(g + 2, (unsigned long long)fq);
                }
                sts_v2_s64(cell, 0, 0);
            }
        };
        int it = 0;
        for (int tile = tile_begin; tile < tile_end; tile += tile_inc, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            int tm, tn;
            decode_tile(tile, tm, tn);
            EpiRow<BN> er;
            int64_t row0, row_end;
            cta_span<CG>(p, tm, rank, row0, row_end);
            epilogue_prepare<BN>(p, row0 + row_in_tile, row_end, tn, lane, er);
            if (p.stats) {                             // a new (image, column tile) run: flush the previous one's sums
                const int img = (int)(row0 / p.tile_rpi);
                if (img != run_img || tn != run_tn) {
                    stats_flush(run_img, run_tn);
                    run_img = img;
                    run_tn = tn;
                }
            }
            mbar_wait_backoff(&acc_full[acc], acc_phase, 100);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16);
            // FMT is a template parameter of the KERNEL: with both formats' epilogues inlined the code was 190 KB and
            // ncu charged 11 % of the samples of a K = 1024 conv to instruction fetch (no_inst)
            epilogue_rows<BN, FMT>(p, taddr, tn, lane, half, er, stats_acc);
            // release this accumulator stage back to the (leader's) MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 1) mbar_arrive(&acc_empty[acc]);
                else mbar_arrive_cluster(mapa_u32(&acc_empty[acc], 0));
            }
        }
        if (p.stats) stats_flush(run_img, run_tn);
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ================================================================================================ swapped form
// Narrow-output convs / linears (N = Cout = 128: the VAE's full-resolution level).  With the pixels on the MMA's M
// side a 128 x 128 tile re-reads 8 KB of operands from shared memory per 64 tensor cycles -- the 128 B/clk
// shared-memory port caps the tensor pipe at ~50 % (measured 750 TFLOP/s).  Here the roles are swapped:
//     D^T[cout, pixel] = W[cout, K] * X[pixel, K]^T      A := weights (M = 128 = Cout),  B := 256 pixels (N = 256)
// which has the operand traffic of the wide-N case (12 KB per 128 cycles).  The nine row-shifted K segments of the
// implicit conv apply to the B operand.  The accumulator comes out channel-major (TMEM lane = output channel,
// column = pixel), so the epilogue turns each 32 channel x 32 pixel block through a warp-private shared-memory tile
// into pixel-major 64-byte row pieces (16-bit NHWC); the residual comes in the same way.  GroupNorm statistics are
// free in this layout: a thread owns ONE channel and just sums its pixels in two registers across all its tiles.
constexpr int TBN = 256;                            // pixels per tile
constexpr int T_STAGE_BYTES = A_STAGE_BYTES + TBN * BLOCK_K * 2;   // weights 16 KB + pixels 32 KB
constexpr int T_PB = 34 * 1024;                     // grouped mode: 256 + 8 pixel rows x 128 B, rounded up to the 1 KB swizzle repeat
// [32 pixels][32 channels] 16-bit, pitch 80 B: the eight 16-byte rows of an stmatrix / ldmatrix 8x8 block (pixels r .. r + 7,
// one 8-channel piece) and the eight rows a quarter-warp moves on the coalesced side (same piece of eight pixels) fall
// in eight different 16-byte bank groups (5 r mod 8).
constexpr int T_STG_PITCH = 80;
constexpr int T_STG_WARP = 32 * T_STG_PITCH;
constexpr int T_STAGING = EPI_WARPS * T_STG_WARP;

template <int FMT>
__global__ void __launch_bounds__(NUM_THREADS, 1) smtl_gemmT_kernel(const __grid_constant__ GemmKParams p) {
    constexpr int TMEM_COLS = 512;
    constexpr int ACC_STRIDE = 256;
    constexpr int MAX_STAGES = 12;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int stages = p.stages;
    const bool grouped = p.grouped != 0;
    uint8_t* smem_w = smem + (size_t)p.sp * T_PB;                     // grouped: weight ring after the pixel ring
    uint8_t* ring_end = grouped ? smem_w + (size_t)p.sw * A_STAGE_BYTES : smem + (size_t)stages * T_STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring_end);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + MAX_STAGES;
    uint64_t* acc_full = bars + 2 * MAX_STAGES;
    uint64_t* acc_empty = bars + 2 * MAX_STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
    uint8_t* staging = ring_end + 512;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.tm_a0);
        tma_prefetch_desc(&p.tm_a1);
        tma_prefetch_desc(&p.tm_b);
        for (int s = 0; s < MAX_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], EPI_WARPS); }
        fence_mbar_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int num_tiles = p.tiles_m;                 // 256-pixel blocks

    if (warp == 0) {
        if (grouped) {
            int ps = 0, ws = 0;
            uint32_t pph = 0, wph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int64_t pix0, pix_end;
                tile_span(p, tile, TBN, pix0, pix_end);
                for (int g = 0; g < p.ngrp; ++g) {
                    const GemmKParams::Grp gr = p.grp[g];
                    const CUtensorMap* tmx = gr.src ? &p.tm_a1 : &p.tm_a0;
                    for (int kb = 0; kb < gr.kblocks; ++kb) {
                        mbar_wait(&empty_bar[ps], pph ^ 1u);
                        uint8_t* sx = smem + (size_t)ps * T_PB;
                        const int32_t xrow = (int32_t)(pix0 + gr.row_shift);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&full_bar[ps], (TBN + 8) * BLOCK_K * 2);
                            tma_load_2d(sx, tmx, &full_bar[ps], gr.a_col0 + kb * BLOCK_K, xrow);                     // 256 rows
                            tma_load_2d(sx + TBN * BLOCK_K * 2, &p.tm_x8[gr.src ? 1 : 0], &full_bar[ps],
                                        gr.a_col0 + kb * BLOCK_K, xrow + TBN);                                         // + 8 rows
                        }
                        __syncwarp();
                        if (++ps == p.sp) { ps = 0; pph ^= 1u; }
                        for (int sub = 0; sub < gr.nsub; ++sub) {
                            const int wb = p.sp + ws;
                            mbar_wait(&empty_bar[wb], wph ^ 1u);
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&full_bar[wb], A_STAGE_BYTES);
                                tma_load_2d(smem_w + (size_t)ws * A_STAGE_BYTES, &p.tm_b, &full_bar[wb],
                                            (gr.kb0 + sub * gr.kblocks + kb) * BLOCK_K, 0);
                            }
                            __syncwarp();
                            if (++ws == p.sw) { ws = 0; wph ^= 1u; }
                        }
                    }
                }
            }
        } else {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int64_t pix0, pix_end;
                tile_span(p, tile, TBN, pix0, pix_end);
                int kb_global = 0;
                for (int s = 0; s < p.nseg; ++s) {
                    const smtl_gemm_seg sg = p.seg[s];
                    const CUtensorMap* tmx = sg.src ? &p.tm_a1 : &p.tm_a0;
                    const int32_t xrow = (int32_t)(pix0 + sg.row_shift);
                    for (int kb = 0; kb < sg.kblocks; ++kb, ++kb_global) {
                        mbar_wait(&empty_bar[stage], phase ^ 1u);
                        uint8_t* sw = smem + (size_t)stage * T_STAGE_BYTES;     // weights  [128 x 64]
                        uint8_t* sx = sw + A_STAGE_BYTES;                        // pixels   [256 x 64]
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&full_bar[stage], T_STAGE_BYTES);
                            tma_load_2d(sw, &p.tm_b, &full_bar[stage], kb_global * BLOCK_K, 0);
                            tma_load_2d(sx, tmx, &full_bar[stage], sg.a_col0 + kb * BLOCK_K, xrow);
                        }
                        __syncwarp();
                        if (++stage == stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // converged warp, elect.sync at the issue (see smtl_gemm_kernel)
        if (grouped) {
            const uint32_t IDESC = make_idesc_16(BLOCK_M, TBN, 0, 0, p.fmt);
            int ps = 0, ws = 0;
            uint32_t pph = 0, wph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * ACC_STRIDE;
                uint32_t accum = 0;
                for (int g = 0; g < p.ngrp; ++g) {
                    const int nsub = p.grp[g].nsub, nkb = p.grp[g].kblocks;
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&full_bar[ps], pph);
                        const uint32_t sx = smem_u32(smem + (size_t)ps * T_PB);
                        for (int sub = 0; sub < nsub; ++sub) {
                            const int wb = p.sp + ws;
                            mbar_wait(&full_bar[wb], wph);
                            MAINLOOP_FENCE();
                            const uint64_t da = make_smem_desc_sw128(smem_u32(smem_w + (size_t)ws * A_STAGE_BYTES));
                            const uint64_t db = make_smem_desc_sw128(sx + sub * 128);            // tap `sub`: one pixel row further
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < BLOCK_K / 16; ++k)
                                    tc_mma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, accum | (uint32_t)k);
                                tc_commit(&empty_bar[wb]);
                            }
                            accum = 1;
                            __syncwarp();
                            if (++ws == p.sw) { ws = 0; wph ^= 1u; }
                        }
                        if (elect_one()) tc_commit(&empty_bar[ps]);
                        __syncwarp();
                        if (++ps == p.sp) { ps = 0; pph ^= 1u; }
                    }
                }
                if (elect_one()) tc_commit(&acc_full[acc]);
                __syncwarp();
            }
        } else {
            const uint32_t IDESC = make_idesc_16(BLOCK_M, TBN, 0, 0, p.fmt);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&acc_empty[acc], ((it >> 1) & 1) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * ACC_STRIDE;
                for (int kb = 0; kb < p.total_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    MAINLOOP_FENCE();
                    const uint32_t sw = smem_u32(smem + (size_t)stage * T_STAGE_BYTES);
                    const uint64_t da = make_smem_desc_sw128(sw);
                    const uint64_t db = make_smem_desc_sw128(sw + A_STAGE_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BLOCK_K / 16; ++k)
                            tc_mma_f16(tmem_d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), IDESC, (kb | k) != 0);
                        tc_commit(&empty_bar[stage]);
                        if (kb == p.total_kb - 1) tc_commit(&acc_full[acc]);
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        // The accumulator is channel-major (TMEM lane = output channel, column = pixel); the output is pixel-major.
        // tcgen05.ld.16x256b hands a thread the MMA-fragment shape (channels g, g + 8 of a 16-lane half x pixel PAIRS), so
        // a packed pair is one stmatrix fragment and stmatrix.trans writes [pixel][8 channels] rows: 4 transposing
        // stores per 32 x 32 block instead of 32 two-byte ones, half the 16-bit conversions, and the residual comes back
        // the same way through ldmatrix.trans.  (The first version -- one channel per thread, 32 st.shared.u16 + 32
        // ld.shared.u16 per block -- shared the 128 B/clk shared-memory port with the MMAs' 96 B/clk of operand reads:
        // tensor pipe 73 % active, 15 M bank conflicts per launch.)
        const int quarter = warp & 3;
        const int half = (warp - 2) >> 2;
        const int g = lane >> 2, tq = lane & 3;
        // statistics: after the quad reduction lane (g, tq) owns channel 32 quarter + 8 tq + g
        const int ch = quarter * 32 + 8 * tq + g;
        const bool ch_ok = ch < p.n;
        float bias_c[4];                                              // channels 32 quarter + 8 m + g, m = 2 h + j
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const int c = quarter * 32 + 8 * m + g;
            bias_c[m] = (p.bias && c < p.n) ? __ldg(p.bias + c) : 0.0f;
        }
        const uint32_t stg = smem_u32(staging + (warp - 2) * T_STG_WARP);      // explicit shared-space accesses below
        // coalesced side: in step i a lane moves the 8-channel piece (lane >> 3) of pixel 8 i + (lane & 7)
        const int piece = lane >> 3;
        const int prow = lane & 7;
        const int cpiece = quarter * 32 + piece * 8;                  // its first channel
        const bool piece_ok = cpiece + 8 <= p.n;
        // matrix side: thread 8 i + r addresses row r (= pixel) of matrix i of an x4 group; group (h, kk) holds the
        // matrices (pixel block 2 kk + (i >> 1), channel block 2 h + (i & 1))
        const uint32_t mat_off = (uint32_t)((8 * (lane >> 4) + (lane & 7)) * T_STG_PITCH + 16 * ((lane >> 3) & 1));
        // this lane's channel sums of the current image, exact: every tile's fp32 partial is added in the int64
        // fixed-point form of the statistics cells (smtl_common.cuh), so the order of the tiles does not matter
        // (the fp32 partial of one TILE -- the same pixels wherever the image sits in the batch -- is what gets fixed)
        long long acc_sl = 0, acc_sh = 0, acc_ql = 0, acc_qh = 0;
        float2 tsum[4], tsq[4];                                       // this tile's running partials of channel m (even, odd pixels)
#pragma unroll
        for (int m = 0; m < 4; ++m) { tsum[m] = make_float2(0.f, 0.f); tsq[m] = make_float2(0.f, 0.f); }
        int cur_img = -1;
        auto commit = [&]() {                                         // tile partials -> the int64 accumulators
            float s_mine = 0.f, q_mine = 0.f;
#pragma unroll
            for (int m = 0; m < 4; ++m) {                             // the four lanes of a quad hold the same channels
                float a = tsum[m].x + tsum[m].y, b = tsq[m].x + tsq[m].y;
                a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
                b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2);
                if (m == tq) { s_mine = a; q_mine = b; }
                tsum[m] = make_float2(0.f, 0.f);
                tsq[m] = make_float2(0.f, 0.f);
            }
            bool hi;
            long long v = stats_fix(s_mine, hi);
            if (hi) acc_sh += v; else acc_sl += v;
            v = stats_fix(q_mine, hi);
            if (hi) acc_qh += v; else acc_ql += v;
        };
        auto flush = [&]() {
            if (p.stats && cur_img >= 0 && ch_ok) {
                unsigned long long* dst =
                    p.stats + (((int64_t)(blockIdx.x % p.stats_rep) * p.stats_images + cur_img) * p.n + ch) * 4;
                if (acc_sl) atomicAdd(dst, (unsigned long long)acc_sl);
                if (acc_sh) atomicAdd(dst + 1, (unsigned long long)acc_sh);
                if (acc_ql) atomicAdd(dst + 2, (unsigned long long)acc_ql);
                if (acc_qh) atomicAdd(dst + 3, (unsigned long long)acc_qh);
            }
            acc_sl = acc_sh = acc_ql = acc_qh = 0;
        };
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            int64_t pix0, pix_end;
            tile_span(p, tile, TBN, pix0, pix_end);
            if (p.stats) {                                            // tiles are image-aligned whenever statistics are produced
                const int img = tile / p.tiles_per_img;
                if (img != cur_img) { flush(); cur_img = img; }
            }
            mbar_wait(&acc_full[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + acc * ACC_STRIDE + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
            for (int c0 = half * 32; c0 < TBN; c0 += 64) {
                // pixel (c0 + lane): output row, validity
                const int64_t grow = pix0 + c0 + lane;
                bool ok, halo;
                int64_t orow;
                int map_img;
                map_row(p, grow, pix_end, ok, orow, halo, map_img);
                halo = halo && (p.rowmap == SMTL_ROWMAP_PAD_KEEP);
                const uint32_t okmask = __ballot_sync(0xffffffffu, ok);
                const uint32_t halomask = __ballot_sync(0xffffffffu, halo);
                if ((okmask | halomask) == 0 && pix0 + c0 >= pix_end) break;         // warp-uniform: past the end
                const int orow32 = (int)orow;
                uint32_t rr[2][16];
                tmem_ld_16x256b_x4(taddr + c0, rr[0]);                               // channels quarter * 32 + 0 .. 15
                tmem_ld_16x256b_x4(taddr + (16u << 16) + c0, rr[1]);                 //                     + 16 .. 31
                // rows this lane moves on the coalesced side: pixel 8 i + prow
                int orow_i[4];
                uint32_t ok_i = 0, halo_i = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    orow_i[i] = __shfl_sync(0xffffffffu, orow32, 8 * i + prow);
                    ok_i |= ((okmask >> (8 * i + prow)) & 1u) << i;
                    halo_i |= ((halomask >> (8 * i + prow)) & 1u) << i;
                }
                uint32_t rp[2][8];                                       // the residual as fragments (pixel pairs)
                if (p.res1) {                                            // 16-bit residual: global -> staging tile -> ldmatrix.trans
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 q = make_uint4(0, 0, 0, 0);
                        if (((ok_i >> i) & 1u) && piece_ok)
                            q = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.res1) +
                                                                     (int64_t)orow_i[i] * p.ldres + cpiece));
                        sts_v4(stg + (8 * i + prow) * T_STG_PITCH + piece * 16, q);
                    }
                    __syncwarp();
#pragma unroll
                    for (int h = 0; h < 2; ++h)
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk)
                            ldmatrix_x4_trans(stg + mat_off + (uint32_t)(16 * kk * T_STG_PITCH + 32 * h), rp[h][4 * kk],
                                              rp[h][4 * kk + 1], rp[h][4 * kk + 2], rp[h][4 * kk + 3]);
                    __syncwarp();
                }
                tmem_ld_wait();
                uint32_t w[2][8];                                        // w[h][2 k + j]: pixels 8 k + 2 tq + {0, 1}, channel 16 h + 8 j + g
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int m = 2 * h + j;
                            float v0 = __uint_as_float(rr[h][4 * k + 2 * j]) + bias_c[m];
                            float v1 = __uint_as_float(rr[h][4 * k + 2 * j + 1]) + bias_c[m];
                            if (p.res1) {
                                const float2 f = unpack16x2(rp[h][2 * k + j], FMT);
                                v0 += f.x;
                                v1 += f.y;
                            }
                            w[h][2 * k + j] = pack16x2(v0, v1, FMT);
                            if (p.stats) {                               // packed: one FADD2 + one FFMA2 per pixel pair
                                if (okmask != 0xffffffffu) {                 // warp-uniform; halo / out-of-range pixels count as 0
                                    const int px = 8 * k + 2 * tq;
                                    v0 = ((okmask >> px) & 1u) ? v0 : 0.f;
                                    v1 = ((okmask >> (px + 1)) & 1u) ? v1 : 0.f;
                                }
                                tsum[m] = fadd2(tsum[m], make_float2(v0, v1));
                                tsq[m] = ffma2(make_float2(v0, v1), make_float2(v0, v1), tsq[m]);
                            }
                        }
                    }
                }
                // channel-major -> pixel-major: four transposing 8x8 stores per 16-channel half, then 64-byte row pieces
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk)
                        stmatrix_x4_trans(stg + mat_off + (uint32_t)(16 * kk * T_STG_PITCH + 32 * h), w[h][4 * kk], w[h][4 * kk + 1],
                                          w[h][4 * kk + 2], w[h][4 * kk + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint4 q = lds_v4(stg + (8 * i + prow) * T_STG_PITCH + piece * 16);
                    if (((ok_i >> i) & 1u) && piece_ok)
                        *reinterpret_cast<uint4*>(p.out_bf16 + (int64_t)orow_i[i] * p.ldc + cpiece) = q;
                    else if (((halo_i >> i) & 1u) && piece_ok)       // PAD_KEEP: zero halo
                        *reinterpret_cast<uint4*>(p.out_bf16 + (int64_t)orow_i[i] * p.ldc + cpiece) = make_uint4(0, 0, 0, 0);
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            if (p.stats) commit();
        }
        flush();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int BN, int CG, int FMT>
int launch_gemm_fmt(const GemmKParams& kp, int grid, int smem_bytes, cudaStream_t stream) {
    static std::atomic<uint64_t> attr_devs{0};   // per instantiation, one bit per device ordinal
    if (smtl_host::first_use_on_device(attr_devs))
        SMTL_CHECK_CUDA(cudaFuncSetAttribute(smtl_gemm_kernel<BN, CG, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             SMEM_BUDGET));
    if (CG == 1) {
        smtl_gemm_kernel<BN, CG, FMT><<<grid, NUM_THREADS, smem_bytes, stream>>>(kp);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid, 1, 1);
        cfg.blockDim = dim3(NUM_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        SMTL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, smtl_gemm_kernel<BN, CG, FMT>, kp));
    }
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}
template <int BN, int CG>
int launch_gemm(const GemmKParams& kp, int grid, int smem_bytes, cudaStream_t stream) {
#ifdef SMTL_GEMM_BF16_PART
    return launch_gemm_fmt<BN, CG, FMT_BF16>(kp, grid, smem_bytes, stream);
#else
    return kp.fmt == FMT_F16 ? launch_gemm_fmt<BN, CG, FMT_F16>(kp, grid, smem_bytes, stream)
                             : smtl_detail_gemm_launch_bf16(BN, CG, &kp, grid, smem_bytes, stream);
#endif
}

template <int BN>
int launch_gemm_cg(const GemmKParams& kp, int cg, int grid, int smem_bytes, cudaStream_t stream) {
    return cg == 2 ? launch_gemm<BN, 2>(kp, grid, smem_bytes, stream) : launch_gemm<BN, 1>(kp, grid, smem_bytes, stream);
}

int pick_block_n(int n, int act, int total_kb) {
    if (act == SMTL_ACT_GEGLU) return 256;
    if (n <= 32) return 32;
    if (n <= 64) return 64;
    if (n <= 128) return 128;
    const int cands[] = {128, 160, 192, 224, 256};
    int best = 256;
    double best_cost = 1e30;
    for (int bn : cands) {
        const int tiles = (n + bn - 1) / bn;
        double cost;
        if (total_kb >= 16) {
            // long-K (convs): a k-block costs max(bn / 2 tensor cycles, ~550 clk of barrier hand-offs), so up to ~270
            // columns are free: fewest tiles wins, then fewest padded columns (N = 640: 3 x 224 beats 4 x 160 by 18 %)
            cost = (double)tiles * (bn > 275 ? bn : 275) + 1e-3 * tiles * bn;
        } else {
            // short-K linears are epilogue / store bound: MMA time ~ tiles * max(bn, 160), small bonus for wide tiles
            cost = (double)tiles * (bn > 160 ? bn : 160) * (1.0 + 16.0 / bn);
        }
        if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
    }
    return best;
}

}  // namespace

#ifdef SMTL_GEMM_BF16_PART
int smtl_detail_gemm_launch_bf16(int block_n, int cg, const void* kpv, int grid, int smem_bytes, cudaStream_t st) {
    const GemmKParams& kp = *reinterpret_cast<const GemmKParams*>(kpv);
    if (cg == 3) {
        static std::atomic<uint64_t> attr_devs{0};
        if (smtl_host::first_use_on_device(attr_devs))
            SMTL_CHECK_CUDA(cudaFuncSetAttribute(smtl_gemmT_kernel<FMT_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
        smtl_gemmT_kernel<FMT_BF16><<<grid, NUM_THREADS, smem_bytes, st>>>(kp);
        SMTL_CHECK_CUDA(cudaGetLastError());
        return SMTL_OK;
    }
    switch (block_n) {
        case 32: return launch_gemm_cg<32>(kp, cg, grid, smem_bytes, st);
        case 64: return launch_gemm_cg<64>(kp, cg, grid, smem_bytes, st);
        case 128: return launch_gemm_cg<128>(kp, cg, grid, smem_bytes, st);
        case 160: return launch_gemm_cg<160>(kp, cg, grid, smem_bytes, st);
        case 192: return launch_gemm_cg<192>(kp, cg, grid, smem_bytes, st);
        case 224: return launch_gemm_cg<224>(kp, cg, grid, smem_bytes, st);
        case 256: return launch_gemm_cg<256>(kp, cg, grid, smem_bytes, st);
        default: smtl_host::set_error("gemm_run: bad block_n %d", block_n); return SMTL_EINVAL;
    }
}
#else
extern "C" int smtl_gemm_plan(const smtl_gemm_args* a, smtl_gemm_op* op) {
    SMTL_CHECK_ARG(a && op, "gemm_plan: NULL argument");
    SMTL_CHECK_ARG(a->a0 && a->b, "gemm_plan: NULL operand");
    SMTL_CHECK_ARG(a->m > 0 && a->n > 0 && a->k > 0, "gemm_plan: empty problem m=%lld n=%d k=%d", (long long)a->m,
                   a->n, a->k);
    SMTL_CHECK_ARG(a->out_f32 || a->out_bf16 || a->aux_bf16, "gemm_plan: no output");
    SMTL_CHECK_ARG(a->nseg >= 0 && a->nseg <= SMTL_MAX_SEG, "gemm_plan: nseg %d out of range", a->nseg);
    memset(op, 0, sizeof(*op));
    op->args = *a;
    smtl_gemm_args& g = op->args;
    if (g.nseg == 0) {
        g.nseg = 1;
        g.seg[0].row_shift = 0;
        g.seg[0].kblocks = (g.k + BLOCK_K - 1) / BLOCK_K;
        g.seg[0].src = 0;
        g.seg[0].a_col0 = 0;
    }
    int total_kb = 0;
    for (int s = 0; s < g.nseg; ++s) {
        SMTL_CHECK_ARG(g.seg[s].kblocks > 0, "gemm_plan: segment %d has no K blocks", s);
        SMTL_CHECK_ARG(g.seg[s].src == 0 || (g.seg[s].src == 1 && g.a1), "gemm_plan: segment %d bad source", s);
        total_kb += g.seg[s].kblocks;
    }
    SMTL_CHECK_ARG(total_kb * BLOCK_K >= g.k && (total_kb - 1) * BLOCK_K < g.k,
                   "gemm_plan: segments cover %d K blocks but k=%d", total_kb, g.k);
    int bn = g.block_n ? g.block_n : pick_block_n(g.n, g.act, total_kb);
    SMTL_CHECK_ARG(bn == 32 || bn == 64 || bn == 128 || bn == 160 || bn == 192 || bn == 224 || bn == 256,
                   "gemm_plan: unsupported block_n %d", bn);
    if (g.act == SMTL_ACT_GEGLU) {
        SMTL_CHECK_ARG(bn == 256 && g.n % 256 == 0, "gemm_plan: GEGLU needs n %% 256 == 0 (n=%d)", g.n);
        SMTL_CHECK_ARG(!g.bias_per_row, "gemm_plan: GEGLU with per-row bias");
    }
    SMTL_CHECK_ARG(g.fmt16 == SMTL_FMT_BF16 || g.fmt16 == SMTL_FMT_F16, "gemm_plan: bad fmt16 %d", g.fmt16);
    if (g.rowmap != SMTL_ROWMAP_IDENTITY)
        SMTL_CHECK_ARG(g.img_h > 0 && g.img_w > 0, "gemm_plan: conv row map needs img_h/img_w");
    SMTL_CHECK_ARG(g.rowmap >= SMTL_ROWMAP_IDENTITY && g.rowmap <= SMTL_ROWMAP_UP2_PAD, "gemm_plan: bad rowmap");
    SMTL_CHECK_ARG(g.up_parity >= 0 && g.up_parity <= 3, "gemm_plan: bad up_parity");
    SMTL_CHECK_ARG(g.tile_order >= 0 && g.tile_order <= 2, "gemm_plan: bad tile_order");
    if (g.group_rows) {
        SMTL_CHECK_ARG(g.group_rows > 0 && g.m % g.group_rows == 0, "gemm_plan: group_rows %lld must divide m",
                       (long long)g.group_rows);
        SMTL_CHECK_ARG(g.rowmap == SMTL_ROWMAP_IDENTITY && !g.bias_per_row && g.act != SMTL_ACT_GEGLU && !g.stats,
                       "gemm_plan: grouped GEMM is a plain token linear");
    }
    SMTL_CHECK_ARG(g.res_fmt16 == 0 || g.res_fmt16 == 1, "gemm_plan: bad res_fmt16 %d", g.res_fmt16);
    if (g.stats) {
        SMTL_CHECK_ARG(g.stats_rows_per_image > 0 && g.stats_images > 0 && g.stats_replicas >= 1 &&
                           g.stats_replicas <= 64,
                       "gemm_plan: stats needs rows_per_image/images/replicas (got %d/%d/%d)", g.stats_rows_per_image,
                       g.stats_images, g.stats_replicas);
        SMTL_CHECK_ARG(g.act != SMTL_ACT_GEGLU && !g.bias_per_row, "gemm_plan: stats with GEGLU / per-row bias");
        SMTL_CHECK_ARG(g.stats_group == 0 || g.stats_group == 1 || g.stats_group == 2 || g.stats_group == 4 || g.stats_group == 8,
                       "gemm_plan: stats_group=%d (0/1, 2, 4 or 8)", g.stats_group);
        SMTL_CHECK_ARG(g.stats_group <= 1 || g.n % g.stats_group == 0, "gemm_plan: n=%d is not a multiple of stats_group=%d", g.n,
                       g.stats_group);
    }
    SMTL_CHECK_ARG(g.m + 4096 < (int64_t)1 << 31, "gemm_plan: m too large for 32-bit TMA coordinates");
    // image-aligned M tiling whenever statistics are produced (see GemmKParams::tile_rpi)
    int64_t tile_rpi = 0;
    if (g.stats) {
        if (g.rowmap == SMTL_ROWMAP_IDENTITY) tile_rpi = g.stats_rows_per_image;
        else if (g.rowmap == SMTL_ROWMAP_TO_PAD) tile_rpi = (int64_t)g.img_h * g.img_w;
        else tile_rpi = (int64_t)(g.img_h + 2) * (g.img_w + 2);
        SMTL_CHECK_ARG(g.m == tile_rpi * g.stats_images, "gemm_plan: stats: m=%lld is not stats_images=%d x %lld rows",
                       (long long)g.m, g.stats_images, (long long)tile_rpi);
    }
    // grouped GEMM whose groups are not whole tiles: M tiles restart at every group as well (rows of the last tile of a
    // group that belong to the next one are computed with the wrong weights and masked by the tile's row_end)
    if (g.group_rows && g.group_rows % BLOCK_M != 0) tile_rpi = g.group_rows;
    op->tile_rpi = tile_rpi;
    const int tile_units = tile_rpi ? (int)(g.m / tile_rpi) : 0;      // images / weight groups the M tiles restart at
    const int stats_smem = g.stats ? STATS_SMEM : 0;   // running column sums of the epilogue warps (unused by the swapped kernel)
    auto count_tiles_m = [&](int tile_rows) {
        if (!tile_rpi) { op->tiles_per_img = 0; return (int)((g.m + tile_rows - 1) / tile_rows); }
        op->tiles_per_img = (int)((tile_rpi + tile_rows - 1) / tile_rows);
        return op->tiles_per_img * tile_units;
    };

    const int sms = smtl_host::num_sms();
    // ---- shift groups: consecutive-row-shift segments (kx = -1, 0, +1 of one ky) share one activation tile
    {
        const bool allow = true;
        int ng = 0, kb0 = 0;
        bool any = false;
        for (int si = 0; si < g.nseg;) {
            int nsub = 1;
            while (allow && nsub < 3 && si + nsub < g.nseg &&
                   g.seg[si + nsub].row_shift == g.seg[si].row_shift + nsub && g.seg[si + nsub].kblocks == g.seg[si].kblocks &&
                   g.seg[si + nsub].src == g.seg[si].src && g.seg[si + nsub].a_col0 == g.seg[si].a_col0)
                ++nsub;
            op->grp[ng].row_shift = g.seg[si].row_shift;
            op->grp[ng].nsub = nsub;
            op->grp[ng].kblocks = g.seg[si].kblocks;
            op->grp[ng].src = g.seg[si].src;
            op->grp[ng].a_col0 = g.seg[si].a_col0;
            op->grp[ng].kb0 = kb0;
            kb0 += nsub * g.seg[si].kblocks;
            any = any || nsub > 1;
            si += nsub;
            ++ng;
        }
        op->ngrp = ng;
        op->grouped = any ? 1 : 0;
    }
    // swapped form (weights on the MMA's M side) for narrow outputs: N <= 128, plain 16-bit output
    {
        const bool eligible = g.group_rows == 0 && g.n <= 128 && g.n >= 64 && g.n % 8 == 0 && g.act == SMTL_ACT_NONE && !g.bias_per_row &&
                              g.out_bf16 && !g.out_f32 && !g.aux_bf16 && !g.res2 && (!g.res1 || g.res_fmt16 == 1) &&
                              (g.ldc % 8) == 0 && (!g.res1 || (g.ldres % 8) == 0) && g.block_n == 0 &&
                              g.cta_group == 0 && g.m >= (int64_t)sms * TBN;
        if (eligible) {
            op->cta_group = 3;                      // marks the swapped kernel
            op->block_n = TBN;
            op->tiles_m = count_tiles_m(TBN);
            op->tiles_n = 1;
            op->total_kblocks = total_kb;
            int stages = (SMEM_BUDGET - 1024 - 512 - T_STAGING) / T_STAGE_BYTES;
            op->smem_bytes = 1024 + stages * T_STAGE_BYTES + 512 + T_STAGING;
            if (op->grouped) {
                op->sp = 3;
                op->sw = (SMEM_BUDGET - 1024 - 512 - T_STAGING - op->sp * T_PB) / A_STAGE_BYTES;
                if (op->sw > 8) op->sw = 8;
                op->smem_bytes = 1024 + op->sp * T_PB + op->sw * A_STAGE_BYTES + 512 + T_STAGING;
                int rc8 = smtl_host::encode_tmap_bf16_2d(op->tmap_x8[0], g.a0, (uint64_t)g.a0_rows, (uint64_t)g.a0_cols,
                                                         (uint64_t)g.a0_ld, 8);
                if (rc8) return rc8;
                if (g.a1) {
                    rc8 = smtl_host::encode_tmap_bf16_2d(op->tmap_x8[1], g.a1, (uint64_t)g.a1_rows, (uint64_t)g.a1_cols,
                                                         (uint64_t)g.a1_ld, 8);
                    if (rc8) return rc8;
                } else {
                    memcpy(op->tmap_x8[1], op->tmap_x8[0], sizeof(op->tmap_x8[0]));
                }
            }
            op->grid = op->tiles_m < sms ? op->tiles_m : sms;
            int rc = smtl_host::encode_tmap_bf16_2d(op->tmap_a0, g.a0, (uint64_t)g.a0_rows, (uint64_t)g.a0_cols,
                                                    (uint64_t)g.a0_ld, TBN);
            if (rc) return rc;
            if (g.a1) {
                rc = smtl_host::encode_tmap_bf16_2d(op->tmap_a1, g.a1, (uint64_t)g.a1_rows, (uint64_t)g.a1_cols,
                                                    (uint64_t)g.a1_ld, TBN);
                if (rc) return rc;
            } else {
                memcpy(op->tmap_a1, op->tmap_a0, sizeof(op->tmap_a0));
            }
            return smtl_host::encode_tmap_bf16_2d(op->tmap_b, g.b, (uint64_t)g.n, (uint64_t)g.k, (uint64_t)g.ldb,
                                                  BLOCK_M);
        }
    }
    // CTA pairs (tcgen05 cta_group::2, 256-row tiles) whenever there are enough rows to fill the machine with them
    int cg = g.cta_group;
    if (cg == 0) {
        const long long tiles256 = (long long)count_tiles_m(2 * BLOCK_M) * (long long)((g.n + bn - 1) / bn);
        // measured on B200 (scripts/bench_kernels.py gemm): pairs win on long plain-K GEMMs (8192^3: 1170 -> 1295
        // TFLOP/s) and lose on short-K linears, whose cost is the epilogue
        cg = (tiles256 >= sms / 2 && g.nseg == 1 && g.k >= 2048 && g.n >= 256) ? 2 : 1;
        // shift-grouped convs: pairs halve the weight traffic (L2 -> smem and smem -> MMA): +8 % at bn = 256, +2-3 % at 160
        if (op->grouped && bn % 16 == 0 && bn >= 128 && tiles256 >= sms / 2) cg = 2;
        // image-aligned tiles (statistics producers): 256-row pair tiles pad a small map far more than 128-row tiles do
        // (15x20: 374 padded rows = 2 x 256 but 3 x 128).  Measured (scripts/bench_kernels.py stats): the pair is worth
        // ~17 % on these convs -- 30x40 maps (6 x 256 vs 11 x 128 rows, 9 % more padding) still run 8 % faster as pairs,
        // 15x20 maps (33 % more padding) 14 % slower.
        if (g.group_rows) cg = 1;
    }
    SMTL_CHECK_ARG(cg == 1 || cg == 2, "gemm_plan: cta_group %d", cg);
    SMTL_CHECK_ARG(!(g.group_rows && cg == 2), "gemm_plan: a grouped GEMM runs on single CTAs (cta_group 0 or 1)");
    SMTL_CHECK_ARG(cg == 1 || bn % 16 == 0, "gemm_plan: cta_group 2 needs block_n %% 16 == 0");
    // Image-aligned tiles of SMALL maps under a pair: 256-row tiles pad them badly, so the two CTAs take two consecutive
    // 128-row tiles of the 128-row tiling instead (GemmKParams::pair_split) -- the padding of single CTAs with the weight
    // traffic of pairs.  (Before: such convs fell back to single CTAs, measured 14 % slower than an unpadded pair.)
    op->pair_split = 0;
    if (cg == 2 && tile_rpi) {
        const double e2 = (double)tile_rpi / (double)(((tile_rpi + 255) / 256) * 256);
        const double e1 = (double)tile_rpi / (double)(((tile_rpi + 127) / 128) * 128);
        if (e2 * 1.04 < e1) op->pair_split = 1;
    }
    op->cta_group = cg;
    op->block_n = bn;
    if (op->pair_split) {
        const int t128 = count_tiles_m(BLOCK_M);          // also sets tiles_per_img in 128-row units
        op->tiles_m = (t128 + 1) / 2;
    } else {
        op->tiles_m = count_tiles_m(cg * BLOCK_M);
    }
    op->tiles_n = (g.n + bn - 1) / bn;
    op->total_kblocks = total_kb;
    const int stage_bytes = A_STAGE_BYTES + (bn / cg) * BLOCK_K * 2;
    int stages = (SMEM_BUDGET - 1024 - 512 - stats_smem) / stage_bytes;
    if (stages > 8) stages = 8;
    op->smem_bytes = 1024 + stages * stage_bytes + 512 + stats_smem;
    const int a_box_rows = op->grouped ? BLOCK_M + 8 : BLOCK_M;
    if (op->grouped) {
        const int pb = KPB * (BLOCK_M + 8) * BLOCK_K * 2, wb = KPB * (bn / cg) * BLOCK_K * 2;   // KPB k-blocks per stage
        // narrow tiles are bound by the activation stream: give it the deeper ring.  The producer issues the loads in
        // program order, so the weight ring must hold the weights of every activation stage in flight (3 taps each)
        // or it, not the activation ring, sets the prefetch distance: a Cout = 3 conv (BN = 32, 36 clk per MMA) was
        // load-latency bound with 6 weight stages.
        const int sp_max = bn <= 64 ? 8 : 6;
        const int budget = SMEM_BUDGET - 1024 - 512 - stats_smem;
        op->sp = (budget - 6 * wb) / pb;
        if (op->sp < 3) op->sp = 3;
        if (op->sp > sp_max) op->sp = sp_max;
        // wide tiles: the weight ring is the critical stream (3 weight stages per activation stage), so it gets the
        // shared memory -- 2 activation stages + 4 weight stages beat 3 + 3 by 1-2 %, 4 + 2 loses 6 % (bn = 256 pairs)
        if (bn >= 128) op->sp = 2;
        op->sw = (budget - op->sp * pb) / wb;
        while (op->sw < 2 && op->sp > 2) {          // widest tiles in one CTA: trade an activation stage for a weight stage
            --op->sp;
            op->sw = (budget - op->sp * pb) / wb;
        }
        SMTL_CHECK_ARG(op->sw >= 1, "gemm_plan: no room for a weight stage (block_n %d)", bn);
        if (op->sw > 3 * op->sp) op->sw = 3 * op->sp;
        if (op->sw > GEMM_MAX_STAGES - op->sp) op->sw = GEMM_MAX_STAGES - op->sp;
        op->smem_bytes = 1024 + op->sp * pb + op->sw * wb + 512 + stats_smem;
    }
    const long long tiles = (long long)op->tiles_m * op->tiles_n;
    const int slots = sms / cg;                     // CTAs (cg = 1) or CTA pairs (cg = 2) resident at once
    op->grid = (int)(tiles < slots ? tiles : slots) * cg;

    int rc = smtl_host::encode_tmap_bf16_2d(op->tmap_a0, g.a0, (uint64_t)g.a0_rows, (uint64_t)g.a0_cols,
                                            (uint64_t)g.a0_ld, a_box_rows);
    if (rc) return rc;
    if (g.a1) {
        rc = smtl_host::encode_tmap_bf16_2d(op->tmap_a1, g.a1, (uint64_t)g.a1_rows, (uint64_t)g.a1_cols,
                                            (uint64_t)g.a1_ld, a_box_rows);
        if (rc) return rc;
    } else {
        memcpy(op->tmap_a1, op->tmap_a0, sizeof(op->tmap_a0));
    }
    const uint64_t b_rows = g.group_rows ? (uint64_t)(g.m / g.group_rows) * (uint64_t)g.n : (uint64_t)g.n;
    rc = smtl_host::encode_tmap_bf16_2d(op->tmap_b, g.b, b_rows, (uint64_t)g.k, (uint64_t)g.ldb, (uint32_t)(bn / cg));
    return rc;
}

extern "C" int smtl_gemm_run(const smtl_gemm_op* op, void* stream) {
    SMTL_CHECK_ARG(op, "gemm_run: NULL op");
    const smtl_gemm_args& g = op->args;
    GemmKParams kp;
    memcpy(&kp.tm_a0, op->tmap_a0, 128);
    memcpy(&kp.tm_a1, op->tmap_a1, 128);
    memcpy(&kp.tm_b, op->tmap_b, 128);
    kp.m = g.m;
    kp.n = g.n;
    kp.n_out = (g.act == SMTL_ACT_GEGLU) ? g.n / 2 : g.n;
    kp.tiles_m = op->tiles_m;
    kp.tiles_n = op->tiles_n;
    kp.total_kb = op->total_kblocks;
    const int cg = op->cta_group == 2 ? 2 : 1;
    const int stage_bytes = A_STAGE_BYTES + (op->block_n / cg) * BLOCK_K * 2;
    kp.stages = (op->smem_bytes - 1024 - 512 - (g.stats ? STATS_SMEM : 0)) / stage_bytes;
    if (op->cta_group == 3) kp.stages = (op->smem_bytes - 1024 - 512 - T_STAGING) / T_STAGE_BYTES;
    kp.nseg = g.nseg;
    for (int s = 0; s < SMTL_MAX_SEG; ++s) kp.seg[s] = g.seg[s];
    kp.bias = g.bias;
    kp.bias_per_row = g.bias_per_row;
    kp.act = g.act;
    kp.res1 = g.res1;
    kp.res2 = g.res2;
    kp.ldres = g.ldres;
    kp.res16 = g.res_fmt16;
    kp.stats = reinterpret_cast<unsigned long long*>(g.stats);
    kp.tile_rpi = op->tile_rpi;
    kp.pair_split = op->pair_split;
    kp.tiles_per_img = op->tiles_per_img;
    // auto: the contiguous order pays when an (image, column tile) run is long (>= 16 tiles: +1.5-2.5 % on the 60x80
    // and larger maps); with a few tiles per image the round-robin order is 4-13 % faster (measured, same script)
    kp.tile_order = g.tile_order ? (g.tile_order == 2) : ((g.stats && op->cta_group != 3 && op->tiles_per_img >= 16) ? 1 : 0);
    kp.stats_rpi = g.stats_rows_per_image;
    kp.stats_images = g.stats_images;
    kp.stats_rep = g.stats_replicas > 0 ? g.stats_replicas : 1;
    kp.stats_g = g.stats_group > 1 ? g.stats_group : 1;
    // whole 32-column chunks and 32-byte aligned rows everywhere: the lean epilogue has no scalar fallbacks
    {
        auto al = [](const void* q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
        const bool plain = g.act == SMTL_ACT_NONE;
        const int n_out = g.act == SMTL_ACT_GEGLU ? g.n / 2 : g.n;
        const int res_ld = g.res_fmt16 ? 16 : 8;
        bool ok = (plain || g.act == SMTL_ACT_GELU || g.act == SMTL_ACT_GEGLU) && !g.bias_per_row &&
                  (g.act == SMTL_ACT_GEGLU ? (n_out % 32 == 0 && g.n % 32 == 0) : g.n % 16 == 0) && (g.out_bf16 || g.out_f32) &&
                  (!g.bias || al(g.bias, 16));
        ok = ok && (!g.out_bf16 || (g.ldc % 16 == 0 && al(g.out_bf16, 32))) && (!g.out_f32 || (g.ldc % 8 == 0 && al(g.out_f32, 32)));
        ok = ok && (!g.aux_bf16 || (g.ld_aux % 16 == 0 && al(g.aux_bf16, 32)));
        ok = ok && (!g.res1 || (g.ldres % res_ld == 0 && al(g.res1, 32))) && (!g.res2 || (g.res1 && !g.res_fmt16 && al(g.res2, 32)));
        ok = ok && (plain || (!g.res1 && !g.res2 && !g.stats));
        kp.lean = ok ? 1 : 0;
    }
    kp.out_f32 = g.out_f32;
    kp.out_bf16 = reinterpret_cast<uint16_t*>(g.out_bf16);
    kp.aux_bf16 = reinterpret_cast<uint16_t*>(g.aux_bf16);
    kp.ldc = g.ldc;
    kp.ld_aux = g.ld_aux;
    kp.group_rows = g.group_rows;
    kp.grouped = op->grouped;
    kp.ngrp = op->ngrp;
    kp.sp = op->sp;
    kp.sw = op->sw;
    for (int i = 0; i < SMTL_MAX_SEG; ++i) {
        kp.grp[i].row_shift = op->grp[i].row_shift; kp.grp[i].nsub = op->grp[i].nsub; kp.grp[i].kblocks = op->grp[i].kblocks;
        kp.grp[i].src = op->grp[i].src; kp.grp[i].a_col0 = op->grp[i].a_col0; kp.grp[i].kb0 = op->grp[i].kb0;
    }
    memcpy(&kp.tm_x8[0], op->tmap_x8[0], 128);
    memcpy(&kp.tm_x8[1], op->tmap_x8[1], 128);
    kp.rowmap = g.rowmap;
    kp.up_py = g.up_parity >> 1;
    kp.up_px = g.up_parity & 1;
    kp.img_h = g.img_h;
    kp.img_w = g.img_w;
    kp.fmt = g.fmt16;
    kp.fd_img = make_fastdiv(1);
    kp.fd_row = make_fastdiv(1);
    if (g.rowmap == SMTL_ROWMAP_TO_PAD) {
        kp.fd_img = make_fastdiv((uint32_t)(g.img_h * g.img_w));
        kp.fd_row = make_fastdiv((uint32_t)g.img_w);
    } else if (g.rowmap != SMTL_ROWMAP_IDENTITY) {
        kp.fd_img = make_fastdiv((uint32_t)((g.img_h + 2) * (g.img_w + 2)));
        kp.fd_row = make_fastdiv((uint32_t)(g.img_w + 2));
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (op->cta_group == 3) {
        if (kp.fmt != FMT_F16) return smtl_detail_gemm_launch_bf16(0, 3, &kp, op->grid, op->smem_bytes, st);
        static std::atomic<uint64_t> attr_devs{0};
        if (smtl_host::first_use_on_device(attr_devs))
            SMTL_CHECK_CUDA(cudaFuncSetAttribute(smtl_gemmT_kernel<FMT_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET));
        smtl_gemmT_kernel<FMT_F16><<<op->grid, NUM_THREADS, op->smem_bytes, st>>>(kp);
        SMTL_CHECK_CUDA(cudaGetLastError());
        return SMTL_OK;
    }
    switch (op->block_n) {
        case 32: return launch_gemm_cg<32>(kp, cg, op->grid, op->smem_bytes, st);
        case 64: return launch_gemm_cg<64>(kp, cg, op->grid, op->smem_bytes, st);
        case 128: return launch_gemm_cg<128>(kp, cg, op->grid, op->smem_bytes, st);
        case 160: return launch_gemm_cg<160>(kp, cg, op->grid, op->smem_bytes, st);
        case 192: return launch_gemm_cg<192>(kp, cg, op->grid, op->smem_bytes, st);
        case 224: return launch_gemm_cg<224>(kp, cg, op->grid, op->smem_bytes, st);
        case 256: return launch_gemm_cg<256>(kp, cg, op->grid, op->smem_bytes, st);
        default: smtl_host::set_error("gemm_run: bad block_n %d", op->block_n); return SMTL_EINVAL;
    }
}
#endif  // SMTL_GEMM_BF16_PART
