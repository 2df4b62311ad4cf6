// smtl_elem.cu -- the HBM-bound kernels of the StableMTL hot path: GroupNorm / LayerNorm, layout producers
// (zero-halo padding, nearest upsample, im2col, UNet-input assembly), small-key attentions and the task-map
// epilogue.  All are coalesced 16-byte-per-thread streaming kernels with warp-shuffle reductions.
#include "smtl_common.cuh"
#include "smtl_host.h"

namespace {
using namespace smtl;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 8 consecutive channels starting at element index `idx` of a fp32 or 16-bit tensor -> fp32 registers
__device__ __forceinline__ void load8(const void* base, int64_t idx, int is16, int fmt, float (&v)[8]) {
    if (is16) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + idx));
        float2 f;
        f = unpack16x2(u.x, fmt); v[0] = f.x; v[1] = f.y;
        f = unpack16x2(u.y, fmt); v[2] = f.x; v[3] = f.y;
        f = unpack16x2(u.z, fmt); v[4] = f.x; v[5] = f.y;
        f = unpack16x2(u.w, fmt); v[6] = f.x; v[7] = f.y;
    } else {
        const float* ptr = reinterpret_cast<const float*>(base) + idx;
        const float4 a = ldg4(ptr), b = ldg4(ptr + 4);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
}
// the raw 16 bytes when the tensor already is 16-bit (pure copies: upsample / im2col / shortcut operand)
__device__ __forceinline__ uint4 load8_as16(const void* base, int64_t idx, int is16, int fmt) {
    if (is16) return __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(base) + idx));
    float v[8];
    load8(base, idx, 0, fmt, v);
    return make_uint4(pack16x2(v[0], v[1], fmt), pack16x2(v[2], v[3], fmt), pack16x2(v[4], v[5], fmt),
                      pack16x2(v[6], v[7], fmt));
}

// ============================================================================================= GroupNorm
// Group mean / rstd of image `b` from the producers' int64 fixed-point cells (smtl_common.cuh): the cells are exact
// whatever the order of the producer's atomics was; a warp per group converts each cell to double and adds them up in
// a FIXED order (lane-strided loop, then the xor tree), so the result is bit-reproducible.
__device__ __forceinline__ void gn_group_stats(const long long* __restrict__ st0, const long long* __restrict__ st1,
                                               int c0, int c1, int replicas, int batch, int b, int groups, double n,
                                               float eps, float* gmean, float* grstd) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;   // whole warps only
    const int cpg = (c0 + c1) / groups;
    const int ncell = cpg * replicas;
    if (warp < nwarps) {
        for (int g = warp; g < groups; g += nwarps) {
            double a0 = 0.0, a2 = 0.0;
            for (int i = lane; i < ncell; i += 32) {
                const int r = i / cpg, c = g * cpg + (i - r * cpg);
                const long long* st = (c < c0) ? st0 + (((int64_t)r * batch + b) * c0 + c) * 4
                                               : st1 + (((int64_t)r * batch + b) * c1 + (c - c0)) * 4;
                const longlong2 u = __ldg(reinterpret_cast<const longlong2*>(st));
                const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(st + 2));
                a0 += stats_value(u.x, u.y);
                a2 += stats_value(v.x, v.y);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            if (lane == 0) {
                const double mean = a0 / n;
                double var = a2 / n - mean * mean;
                if (var < 0.0) var = 0.0;
                gmean[g] = (float)mean;
                grstd[g] = (float)(1.0 / sqrt(var + (double)eps));
            }
        }
    }
}

// GroupNorm apply from producer-side per-(image, channel) statistics (smtl_gemm_args.stats): one streaming pass.
// grid (blocks_per_image, batch), 256 threads; each thread handles 8 channels of one (padded) pixel per step.
// Template parameters are everything that would otherwise be a per-vector run-time choice: ptxas turns such choices
// into predicated instruction pairs, and a predicated-off instruction still takes its issue slot (the run-time format
// alone doubled the conversions of this kernel).  x16: 16-bit input; FMT: 16-bit format; RAW: also emit the
// un-normalised copy; SAME: input and output share the pixel grid (both padded or both compact).
template <bool x16, int FMT, bool RAW, bool SAME>
__global__ void __launch_bounds__(320, 3) gn_apply2_kernel(
    const void* __restrict__ x0, const void* __restrict__ x1, int c0, int c1, const long long* __restrict__ st0,
    const long long* __restrict__ st1, int replicas, int batch, int h, int w, int groups, float eps,
    const float* __restrict__ gamma, const float* __restrict__ beta, int do_silu, int pad_out,
    uint16_t* __restrict__ out, uint16_t* __restrict__ raw, FastDiv div_wp, int in_pad) {
    extern __shared__ float sm[];   // scale[C], shift[C], gmean[groups], grstd[groups]
    const int C = c0 + c1;
    float* scale = sm;
    float* shift = sm + C;
    float* gmean = sm + 2 * C;
    float* grstd = gmean + groups;
    const int b = blockIdx.y;
    const int hw = h * w;
    const int cpg = C / groups;
    gn_group_stats(st0, st1, c0, c1, replicas, batch, b, groups, (double)hw * cpg, eps, gmean, grstd);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float sc = grstd[g] * gamma[c];
        scale[c] = sc;
        shift[c] = beta[c] - gmean[g] * sc;
    }
    __syncthreads();
    // blockDim.x is a multiple of C/8 (host): a thread owns ONE 8-channel vector for the whole kernel, so its
    // scale/shift live in registers and the only index arithmetic left per pixel is pixel -> (y, x).
    const int cv8 = C >> 3;
    const int hp = pad_out ? h + 2 : h, wp = pad_out ? w + 2 : w;
    const uint32_t npix = (uint32_t)hp * wp;                   // per image
    const int ppb = blockDim.x / cv8;                          // pixels one block covers per step
    const int cvi = threadIdx.x % cv8;
    const int pl = threadIdx.x / cv8;
    const int c = cvi * 8;
    // SiLU(y) = y * sigmoid(y) = h * tanh(h) + h with h = y / 2: the halving is folded into (scale, shift), so a
    // normalised + activated element is FFMA, MUFU.TANH, FFMA (this kernel is bound by instruction issue, not by HBM,
    // once the SM clock sits at the power cap: 36 instead of ~70 instructions per 8-channel vector)
    const float pre = (do_silu == 1) ? 0.5f : 1.0f;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = pre * scale[c + i]; sh[i] = pre * shift[c + i]; }
    const void* src;
    int ld, cc;
    if (c < c0) { src = x0; ld = c0; cc = c; } else { src = x1; ld = c1; cc = c - c0; }
    constexpr int U = 4;                                       // independent 16-byte loads in flight per thread
    // A block owns a contiguous run of the image's output pixels and a thread walks it with stride ppb, carrying the
    // output coordinates and pointers along: the per-pixel divisions, 64-bit multiplies and layout branches of the
    // first version were ~100 of its ~160 instructions per 16-byte vector (ncu), and made the kernel issue-bound.
    const int64_t in_base = in_pad ? (int64_t)b * (h + 2) * (w + 2) : (int64_t)b * hw;
    const uint32_t chunk = (npix + gridDim.x - 1) / gridDim.x;
    const uint32_t p_begin = blockIdx.x * chunk;
    const uint32_t p_end = min(npix, p_begin + chunk);
    uint32_t p = p_begin + pl;
    int y = (int)fastdiv(p, div_wp), x = (int)p - y * wp;      // coordinates on the OUTPUT grid (hp x wp)
    int in_w, in_off;                                          // input pixel = y * in_w + x + in_off
    if (SAME) { in_w = wp; in_off = 0; }                                  // same grid: input pixel = output pixel
    else if (pad_out) { in_w = w; in_off = -w - 1; }                      // compact in, padded out
    else { in_w = w + 2; in_off = w + 3; }                                // padded in, compact out
    const char* in_img = reinterpret_cast<const char*>(src) + (in_base * ld + cc) * (x16 ? 2 : 4);
    const int64_t in_pitch = (int64_t)ld * (x16 ? 2 : 4);      // bytes per input pixel
    uint16_t* out_ptr = out + ((int64_t)b * npix + p) * C + c;
    uint16_t* raw_ptr = RAW ? raw + ((int64_t)b * npix + p) * C + c : nullptr;
    const char* in_ptr = in_img + (int64_t)p * in_pitch;       // SAME: walks with the output pointer
    const int64_t in_step = (int64_t)ppb * in_pitch;
    const int64_t out_step = (int64_t)ppb * C;                 // elements between this thread's consecutive pixels
    for (; p < p_end; p += U * ppb, out_ptr += U * out_step, in_ptr += U * in_step) {
        uint4 u[U];
        float4 f0[U], f1[U];
        bool live[U], inter[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            live[k] = p + k * ppb < p_end;
            inter[k] = live[k] && (!pad_out || (y >= 1 && y <= h && x >= 1 && x <= w));
            u[k] = make_uint4(0, 0, 0, 0);
            if (inter[k]) {
                const char* src_px = SAME ? in_ptr + k * in_step : in_img + (int64_t)(y * in_w + x + in_off) * in_pitch;
                if (x16) {
                    u[k] = __ldg(reinterpret_cast<const uint4*>(src_px));
                } else {
                    f0[k] = ldg4(reinterpret_cast<const float*>(src_px));
                    f1[k] = ldg4(reinterpret_cast<const float*>(src_px) + 4);
                }
            }
            x += ppb;
            while (x >= wp) { x -= wp; ++y; }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (!live[k]) continue;
            uint4 o = make_uint4(0, 0, 0, 0), r = make_uint4(0, 0, 0, 0);
            if (inter[k]) {
                float v[8];
                if (x16) {
                    if (RAW) r = u[k];
                    float2 f;
                    f = unpack16x2(u[k].x, FMT); v[0] = f.x; v[1] = f.y;
                    f = unpack16x2(u[k].y, FMT); v[2] = f.x; v[3] = f.y;
                    f = unpack16x2(u[k].z, FMT); v[4] = f.x; v[5] = f.y;
                    f = unpack16x2(u[k].w, FMT); v[6] = f.x; v[7] = f.y;
                } else {
                    v[0] = f0[k].x; v[1] = f0[k].y; v[2] = f0[k].z; v[3] = f0[k].w;
                    v[4] = f1[k].x; v[5] = f1[k].y; v[6] = f1[k].z; v[7] = f1[k].w;
                    if (RAW) {
                        r.x = pack16x2(v[0], v[1], FMT); r.y = pack16x2(v[2], v[3], FMT);
                        r.z = pack16x2(v[4], v[5], FMT); r.w = pack16x2(v[6], v[7], FMT);
                    }
                }
                // packed fp32 pairs (FFMA2): half the issue slots of the two FMAs per element, same results
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    const float2 n = ffma2(make_float2(v[i], v[i + 1]), make_float2(sc[i], sc[i + 1]), make_float2(sh[i], sh[i + 1]));
                    v[i] = n.x; v[i + 1] = n.y;
                }
                if (do_silu == 1) {
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        float2 t;
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(v[i]));
                        asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(v[i + 1]));
                        const float2 h = make_float2(v[i], v[i + 1]);
                        const float2 o2 = ffma2(h, t, h);
                        v[i] = o2.x; v[i + 1] = o2.y;
                    }
                } else if (do_silu == 2) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = silu_exact(v[i]);
                }
                o.x = pack16x2(v[0], v[1], FMT); o.y = pack16x2(v[2], v[3], FMT);
                o.z = pack16x2(v[4], v[5], FMT); o.w = pack16x2(v[6], v[7], FMT);
            }
            *reinterpret_cast<uint4*>(out_ptr + k * out_step) = o;
            if (RAW) *reinterpret_cast<uint4*>(raw_ptr + k * out_step) = r;
        }
        if (RAW) raw_ptr += U * out_step;
    }
}

// stats -> per-(image, channel) (scale, shift); grid = batch, one thread per channel (blockDim = C rounded up to 32)
__global__ void gn_finalize_kernel(const long long* __restrict__ st, int replicas, int batch, int c, int groups, double n_per_group,
                                   float eps, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ ss) {
    extern __shared__ float sm[];   // mean[groups], rstd[groups]
    float* gmean = sm;
    float* grstd = gmean + groups;
    const int b = blockIdx.x;
    const int cpg = c / groups;
    gn_group_stats(st, nullptr, c, 0, replicas, batch, b, groups, n_per_group, eps, gmean, grstd);
    __syncthreads();
    for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
        const int g = ch / cpg;
        const float sc = grstd[g] * gamma[ch];
        reinterpret_cast<float2*>(ss)[(int64_t)b * c + ch] = make_float2(sc, beta[ch] - gmean[g] * sc);
    }
}

// ============================================================================================= LayerNorm
// A warp normalises R rows at once; a lane holds NV float4 of each (C <= 128 NV).  All R * NV loads are issued before
// the first reduction: with one row per warp (and registers sized for C = 1280 whatever C was) the kernel sat at 45 % of
// the copy peak -- too few bytes in flight per SM.
template <bool IN_BF16, int NV, int R, int FMT>
__global__ void ln_kernel(const void* __restrict__ xv, int c, int ldx, int64_t rows, float eps, int64_t rows_per_group,
                          const float* __restrict__ gamma0, const float* __restrict__ beta0,
                          uint16_t* __restrict__ out0, const float* __restrict__ gamma1,
                          const float* __restrict__ beta1, uint16_t* __restrict__ out1, int ldo) {
    const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R;
    if (row0 >= rows) return;
    const int lane = threadIdx.x & 31;
    const int nv = c >> 2;
    float4 v[R][NV];
    float s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s[r] = 0.f;
        const int64_t row = row0 + r;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = lane + 32 * i;
            v[r][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < nv && row < rows) {
                if (IN_BF16) {
                    const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(xv) +
                                                                         row * ldx + 4 * j));
                    const float2 a = unpack16x2(u.x, FMT), b = unpack16x2(u.y, FMT);
                    v[r][i] = make_float4(a.x, a.y, b.x, b.y);
                } else {
                    v[r][i] = ldg4(reinterpret_cast<const float*>(xv) + row * ldx + 4 * j);
                }
            }
        }
    }
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
#pragma unroll
        for (int i = 0; i < NV; ++i) s[r] += v[r][i].x + v[r][i].y + v[r][i].z + v[r][i].w;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] = warp_sum(s[r]) / (float)c;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (lane + 32 * i < nv) {
                const float a = v[r][i].x - mean[r], b = v[r][i].y - mean[r], cc = v[r][i].z - mean[r],
                            d = v[r][i].w - mean[r];
                q += a * a + b * b + cc * cc + d * d;
            }
        }
        s[r] = q;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(warp_sum(s[r]) / (float)c + eps);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t row = row0 + r;
        if (row >= rows) break;
        const int64_t grp = row / rows_per_group;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int j = lane + 32 * i;
            if (j < nv) {
                const float n0 = (v[r][i].x - mean[r]) * rstd[r], n1 = (v[r][i].y - mean[r]) * rstd[r],
                            n2 = (v[r][i].z - mean[r]) * rstd[r], n3 = (v[r][i].w - mean[r]) * rstd[r];
                {
                    const float4 g = ldg4(gamma0 + grp * c + 4 * j), b = ldg4(beta0 + grp * c + 4 * j);
                    uint2 o;
                    o.x = pack16x2(n0 * g.x + b.x, n1 * g.y + b.y, FMT);
                    o.y = pack16x2(n2 * g.z + b.z, n3 * g.w + b.w, FMT);
                    *reinterpret_cast<uint2*>(out0 + row * ldo + 4 * j) = o;
                }
                if (out1) {
                    const float4 g = ldg4(gamma1 + grp * c + 4 * j), b = ldg4(beta1 + grp * c + 4 * j);
                    uint2 o;
                    o.x = pack16x2(n0 * g.x + b.x, n1 * g.y + b.y, FMT);
                    o.y = pack16x2(n2 * g.z + b.z, n3 * g.w + b.w, FMT);
                    *reinterpret_cast<uint2*>(out1 + row * ldo + 4 * j) = o;
                }
            }
        }
    }
}

template <bool IN_BF16, int NV, int R>
static void launch_ln(const smtl_ln_args* a, cudaStream_t st) {
    const int wpb = 8;
    const int64_t rows_per_block = (int64_t)wpb * R;
    const int64_t grid = (a->rows + rows_per_block - 1) / rows_per_block;
    auto kern = a->fmt16 == SMTL_FMT_F16 ? ln_kernel<IN_BF16, NV, R, FMT_F16> : ln_kernel<IN_BF16, NV, R, FMT_BF16>;
    kern<<<(unsigned)grid, wpb * 32, 0, st>>>(a->x, a->c, a->ldx, a->rows, a->eps, a->rows_per_group, a->gamma0, a->beta0,
                                              (uint16_t*)a->out0, a->gamma1, a->beta1, (uint16_t*)a->out1, a->ldo);
}

// ============================================================================================= layout producers
__global__ void upsample_pad_kernel(const void* __restrict__ x, int x16, int batch, int h, int w, int c, int oh, int ow,
                                    uint16_t* __restrict__ out, int fmt) {
    const int cv8 = c >> 3;
    const int hp = oh + 2, wp = ow + 2;
    const int64_t total = (int64_t)batch * hp * wp * cv8;
    const float sy = (float)h / (float)oh, sx = (float)w / (float)ow;   // ATen nearest: src = min(floor(dst*scale), in-1)
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pix = idx / cv8;
        const int cc = (int)(idx - pix * cv8) * 8;
        const int b = (int)(pix / (hp * wp));
        const int rem = (int)(pix - (int64_t)b * hp * wp);
        const int yp = rem / wp, xp = rem - yp * wp;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (yp >= 1 && yp <= oh && xp >= 1 && xp <= ow) {
            const int ysrc = min((int)floorf((float)(yp - 1) * sy), h - 1);
            const int xsrc = min((int)floorf((float)(xp - 1) * sx), w - 1);
            o = load8_as16(x, (((int64_t)b * h + ysrc) * w + xsrc) * c + cc, x16, fmt);
        }
        *reinterpret_cast<uint4*>(out + pix * c + cc) = o;
    }
}

// im2col, 8 channels per thread (c % 8 == 0, kpad == 9*c)
__global__ void im2col_vec8_kernel(const void* __restrict__ x, int x16, int batch, int h, int w, int c, int stride,
                                   int pad_t, int pad_l, int oh, int ow, uint16_t* __restrict__ out, int fmt) {
    const int cv8 = c >> 3;
    const int64_t total = (int64_t)batch * oh * ow * 9 * cv8;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int cc = (int)(idx % cv8) * 8;
        int64_t t = idx / cv8;
        const int tap = (int)(t % 9);
        const int64_t opix = t / 9;
        const int b = (int)(opix / (oh * ow));
        const int rem = (int)(opix - (int64_t)b * oh * ow);
        const int oy = rem / ow, ox = rem - oy * ow;
        const int iy = oy * stride - pad_t + tap / 3, ix = ox * stride - pad_l + tap % 3;
        uint4 o = make_uint4(0, 0, 0, 0);
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
            o = load8_as16(x, (((int64_t)b * h + iy) * w + ix) * c + cc, x16, fmt);
        }
        *reinterpret_cast<uint4*>(out + opix * (int64_t)(9 * c) + tap * c + cc) = o;
    }
}
// im2col, scalar (tiny Cin stems: 3 or 12 channels), zero-fills k in [9c, kpad)
// im2col for tiny Cin stems (3 or 12 channels, fp32 in): one thread produces 8 consecutive k (one 16-byte store) of
// one output pixel; k in [9c, kpad) is zero fill
__global__ void im2col_scalar_kernel(const float* __restrict__ x, int batch, int h, int w, int c, int stride, int pad_t,
                                     int pad_l, int oh, int ow, int kpad, uint16_t* __restrict__ out, int fmt) {
    const int kv = kpad >> 3;
    const int64_t total = (int64_t)batch * oh * ow * kv;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int k0 = (int)(idx % kv) * 8;
        const int64_t opix = idx / kv;
        const int b = (int)(opix / (oh * ow));
        const int rem = (int)(opix - (int64_t)b * oh * ow);
        const int oy = rem / ow, ox = rem - oy * ow;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k0 + i;
            v[i] = 0.f;
            if (k < 9 * c) {
                const int tap = k / c, cc = k - tap * c;
                const int iy = oy * stride - pad_t + tap / 3, ix = ox * stride - pad_l + tap % 3;
                if (iy >= 0 && iy < h && ix >= 0 && ix < w) v[i] = __ldg(x + (((int64_t)b * h + iy) * w + ix) * c + cc);
            }
        }
        *reinterpret_cast<uint4*>(out + opix * kpad + k0) =
            make_uint4(pack16x2(v[0], v[1], fmt), pack16x2(v[2], v[3], fmt), pack16x2(v[4], v[5], fmt),
                       pack16x2(v[6], v[7], fmt));
    }
}

__global__ void rgbprep_kernel(const void* __restrict__ rgbv, int src_u8, int batch, int hw, float* __restrict__ out) {
    const float* rgb = reinterpret_cast<const float*>(rgbv);
    const uint8_t* rgb8 = reinterpret_cast<const uint8_t*>(rgbv);
    const int64_t total = (int64_t)batch * hw;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = idx / hw, p = idx - b * hw;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float v = (src_u8 == 1) ? (float)__ldg(rgb8 + (b * 3 + ch) * hw + p) : __ldg(rgb + (b * 3 + ch) * hw + p);
            out[idx * 3 + ch] = (src_u8 == 2) ? v : v / 255.0f * 2.0f - 1.0f;   // same op order as stablemtl_pipeline.py:263
        }
    }
}

// Stem of the VAE encoder (3 -> C conv, diffusers Encoder.conv_in): [0,255] NCHW rgb -> the 16-bit im2col operand
// [batch*h*w, 64] (k = tap * 3 + channel for the 27 taps x channels, zero fill to 64) in ONE pass: a thread owns one
// pixel, reads its 3x3 neighbourhood from the three planes (neighbouring threads share the loads through L1),
// normalises (stablemtl_pipeline.py:263) and writes 128 contiguous bytes.
__global__ void __launch_bounds__(256) rgb_stem_kernel(const void* __restrict__ rgbv, int src_mode, int batch, int h, int w,
                                                       uint16_t* __restrict__ out, int fmt) {
    const float* rgb = reinterpret_cast<const float*>(rgbv);
    const uint8_t* rgb8 = reinterpret_cast<const uint8_t*>(rgbv);
    const int64_t hw = (int64_t)h * w;
    const int64_t total = (int64_t)batch * hw;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = idx / hw;
        const int p = (int)(idx - b * hw);
        const int y = p / w, x = p - y * w;
        float v[28];
        v[27] = 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
            const bool in = iy >= 0 && iy < h && ix >= 0 && ix < w;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                float t = 0.f;
                if (in) {
                    const int64_t o = (b * 3 + ch) * hw + (int64_t)iy * w + ix;
                    const float raw = (src_mode == 1) ? (float)__ldg(rgb8 + o) : __ldg(rgb + o);
                    t = (src_mode == 2) ? raw : raw / 255.0f * 2.0f - 1.0f;
                }
                v[tap * 3 + ch] = t;
            }
        }
        uint4* dst = reinterpret_cast<uint4*>(out + idx * 64);
        dst[0] = make_uint4(pack16x2(v[0], v[1], fmt), pack16x2(v[2], v[3], fmt), pack16x2(v[4], v[5], fmt), pack16x2(v[6], v[7], fmt));
        dst[1] = make_uint4(pack16x2(v[8], v[9], fmt), pack16x2(v[10], v[11], fmt), pack16x2(v[12], v[13], fmt), pack16x2(v[14], v[15], fmt));
        dst[2] = make_uint4(pack16x2(v[16], v[17], fmt), pack16x2(v[18], v[19], fmt), pack16x2(v[20], v[21], fmt), pack16x2(v[22], v[23], fmt));
        dst[3] = make_uint4(pack16x2(v[24], v[25], fmt), pack16x2(v[26], v[27], fmt), 0u, 0u);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        dst[4] = z; dst[5] = z; dst[6] = z; dst[7] = z;
    }
}

__global__ void unetin_kernel(const float* __restrict__ lat, const int* __restrict__ first_img,
                              const int* __restrict__ second_img, int out_images, int hw, float* __restrict__ out) {
    const int64_t total = (int64_t)out_images * hw;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int img = (int)(idx / hw);
        const int p = (int)(idx - (int64_t)img * hw);
        const float4 a = ldg4(lat + ((int64_t)first_img[img] * hw + p) * 4);
        const float4 b = ldg4(lat + ((int64_t)second_img[img] * hw + p) * 4);
        float4* o = reinterpret_cast<float4*>(out + idx * 12);
        o[0] = a;
        o[1] = b;
        o[2] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ============================================================================================= small attentions
// row softmax fp32 -> bf16 (n <= 8192), one CTA of 256 threads per row
__global__ void softmax_rows_kernel(const float* __restrict__ s, int n, int lds, float scale,
                                    uint16_t* __restrict__ p, int ldp, int fmt) {
    __shared__ float red[8];
    const int64_t row = blockIdx.x;
    const float* src = s + row * lds;
    float v[32];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int j = threadIdx.x + i * 256;
        v[i] = (j < n) ? src[j] * scale : -INFINITY;
        mx = fmaxf(mx, v[i]);
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    __syncthreads();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int j = threadIdx.x + i * 256;
        if (j < n) { v[i] = __expf(v[i] - mx); sum += v[i]; }
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    sum = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += red[i];
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int j = threadIdx.x + i * 256;
        if (j < n) p[row * ldp + j] = to16(v[i] * inv, fmt);
    }
}

// same, 128-bit loads / 64-bit stores: n, lds, ldp multiples of 4 and 16-byte aligned rows; 128 threads per row so
// that more rows are in flight per SM (the kernel is two block reductions deep: latency, not bandwidth)
__global__ void __launch_bounds__(128) softmax_rows_vec4_kernel(const float* __restrict__ s, int n, int lds, float scale,
                                                                uint16_t* __restrict__ p, int ldp, int fmt) {
    __shared__ float red[4];
    const int64_t row = blockIdx.x;
    const float4* src = reinterpret_cast<const float4*>(s + row * lds);
    const int nv = n >> 2;
    float4 v[16];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = threadIdx.x + i * 128;
        if (j < nv) {
            v[i] = __ldg(src + j);
            v[i].x *= scale; v[i].y *= scale; v[i].z *= scale; v[i].w *= scale;
            mx = fmaxf(fmaxf(mx, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
        }
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = threadIdx.x + i * 128;
        if (j < nv) {
            v[i].x = __expf(v[i].x - mx); v[i].y = __expf(v[i].y - mx);
            v[i].z = __expf(v[i].z - mx); v[i].w = __expf(v[i].w - mx);
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
    }
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    const float inv = 1.0f / ((red[0] + red[1]) + (red[2] + red[3]));
    uint2* dst = reinterpret_cast<uint2*>(p + row * ldp);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = threadIdx.x + i * 128;
        if (j < nv) dst[j] = make_uint2(pack16x2(v[i].x * inv, v[i].y * inv, fmt), pack16x2(v[i].z * inv, v[i].w * inv, fmt));
    }
}

struct XattnK {
    int ntok[SMTL_MAX_TASKS];
    int task_of_group[SMTL_MAX_TASKS];
};
// cross-attention on <= 4 constant keys: one warp per token row; a lane owns 8 contiguous channels (one 16-byte
// load), so 8 lanes make a head, a warp covers 4 heads per pass and a q.k dot product closes with 3 shuffles.
__global__ void xattn_kernel(const uint16_t* __restrict__ q, int ldq, int64_t rows, int heads,
                             const float* __restrict__ kc, const float* __restrict__ vc, XattnK tk, int ntp,
                             int64_t rows_per_group, uint16_t* __restrict__ out, int ldo, float scale, int fmt) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const int grp = (int)(row / rows_per_group);
    const int task = tk.task_of_group[grp];
    const int nt = tk.ntok[task];
    const int C = heads * 64;
    const float* kbase = kc + (int64_t)task * ntp * C;      // [ntask, ntp, C], ntp = padded token count (<= 8)
    const float* vbase = vc + (int64_t)task * ntp * C;
    for (int c0 = 0; c0 < C; c0 += 256) {
        const int col = c0 + 8 * lane;
        const bool on = col < C;                       // whole heads: C is a multiple of 64
        float qf[8];
        {
            uint4 u = make_uint4(0, 0, 0, 0);
            if (on) u = __ldg(reinterpret_cast<const uint4*>(q + row * ldq + col));
            float2 t;
            t = unpack16x2(u.x, fmt); qf[0] = t.x; qf[1] = t.y;
            t = unpack16x2(u.y, fmt); qf[2] = t.x; qf[3] = t.y;
            t = unpack16x2(u.z, fmt); qf[4] = t.x; qf[5] = t.y;
            t = unpack16x2(u.w, fmt); qf[6] = t.x; qf[7] = t.y;
        }
        float sc[SMTL_MAX_XATTN_TOKENS];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < SMTL_MAX_XATTN_TOKENS; ++j) {
            float d = 0.f;
            if (j < nt && on) {
                const float4 k0 = ldg4(kbase + j * C + col), k1 = ldg4(kbase + j * C + col + 4);
                d = qf[0] * k0.x + qf[1] * k0.y + qf[2] * k0.z + qf[3] * k0.w + qf[4] * k1.x + qf[5] * k1.y +
                    qf[6] * k1.z + qf[7] * k1.w;
            }
            d += __shfl_xor_sync(0xffffffffu, d, 4);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            sc[j] = (j < nt) ? d * scale : -INFINITY;
            mx = fmaxf(mx, sc[j]);
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < SMTL_MAX_XATTN_TOKENS; ++j) { sc[j] = (j < nt) ? __expf(sc[j] - mx) : 0.f; sum += sc[j]; }
        const float inv = 1.0f / sum;
        float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < SMTL_MAX_XATTN_TOKENS; ++j) {
            if (j < nt && on) {
                const float4 v0 = ldg4(vbase + j * C + col), v1 = ldg4(vbase + j * C + col + 4);
                o[0] += sc[j] * v0.x; o[1] += sc[j] * v0.y; o[2] += sc[j] * v0.z; o[3] += sc[j] * v0.w;
                o[4] += sc[j] * v1.x; o[5] += sc[j] * v1.y; o[6] += sc[j] * v1.z; o[7] += sc[j] * v1.w;
            }
        }
        if (on) {
            uint4 u;
            u.x = pack16x2(o[0] * inv, o[1] * inv, fmt); u.y = pack16x2(o[2] * inv, o[3] * inv, fmt);
            u.z = pack16x2(o[4] * inv, o[5] * inv, fmt); u.w = pack16x2(o[6] * inv, o[7] * inv, fmt);
            *reinterpret_cast<uint4*>(out + row * ldo + col) = u;
        }
    }
}

// ============================================================================================= fused cross-attention
// BasicTransformerBlock's  h += attn2(LN2(h), text);  n3 = LN3(h)  (src/model/attention.py:355-373) in ONE kernel over
// the residual stream.  The prompt is one of a few CONSTANT task names (stablemtl_pipeline.py:464-472), so attn2
// collapses at load time:
//     score[head, j] = (LN2(h) Wq_head^T) . k[j, head] / 8 = xn . (gamma2 * Wq_head^T k[j, head] / 8) + beta2 . (...)
//     attn2(h)       = sum_head sum_j softmax_j(score)[head, j] * (Wo[:, head] v[j, head]) + bo
// i.e. two SKINNY GEMMs per row block -- [16 x C] x [C x V] and [16 x V] x [V x C] with V = heads * n_tok = 20..80
// vectors -- instead of two C x C GEMMs, with xn = (h - mean) * rstd.  The row is read from HBM once (fp32), written once
// (fp32) and emitted once as the 16-bit LN3 operand of the feed-forward: 10 B per element instead of 32 B over five
// kernels.  N = 24..48 is far below a tcgen05 tile, so the contractions run on mma.sync.m16n8k16 (fp32 accumulate)
// straight from registers: a warp owns 16 rows and makes four passes over them (LN2 statistics; scores; output + LN3
// statistics; LN3), the re-reads hitting L1 / L2.  (The first version of this kernel did the dot products on the FMA
// pipe: 7.4 k instructions per 8 rows, 183-206 registers, issue-bound at 3x the HBM floor.)
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, int fmt) {
    if (fmt == FMT_F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// H = heads (C = 64 H), NT = padded token count (4 or 8).  Vectors v = head * NT + token, padded to VP (multiple of 16).
// R16: a warp owns 16 rows (both row halves of the MMA tile) or 8 (the upper half of the tile is fed zeros).
//
// The warp's rows live in REGISTERS for the whole kernel -- a lane holds, of each of its one or two rows, the four
// consecutive channels 4tq .. 4tq+3 of every 16-channel block: 160 fp32 for 16 rows of 320 channels or 8 rows of 640 --
// so HBM is read once (all of a lane's 16-byte loads are issued back to back: 20 KB in flight per warp) and written once.
// (The first mma.sync version re-read its rows from L1 / L2 in passes 2-4; with 24 warps x 20 KB per SM the re-reads
// missed L1 and the kernel sat at 19 % of the DRAM peak on load latency: ncu long_scoreboard on every first use.)
__device__ __forceinline__ float4 ld_stream4(const float* p) {        // read-once data: keep it out of L1 (the tables live there)
    float4 v;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
template <int H, int NT, int FMT, bool R16>
__global__ void __launch_bounds__(128, 2) xattn_mma_kernel(
    float* __restrict__ hs, int ldh, int64_t rows_per_group, int ngroups, XattnK tk, const uint16_t* __restrict__ ap,
    const float* __restrict__ ca, const uint16_t* __restrict__ bmt, const float* __restrict__ bo,
    const float* __restrict__ g3, const float* __restrict__ b3, uint16_t* __restrict__ out, int ldo, float eps2,
    float eps3) {
    constexpr int C = 64 * H, V = H * NT, VP = (V + 15) / 16 * 16, NTILE = VP / 8, KV = VP / 16, KC = C / 16;
    constexpr int R = R16 ? 16 : 8;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, tq = lane & 3;                                // fragment row within 8, column pair
    const int64_t wpg = (rows_per_group + R - 1) / R;                      // warps per row group
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int grp = (int)(gw / wpg);
    if (grp >= ngroups) return;
    const int64_t local0 = (gw - (int64_t)grp * wpg) * R;
    const int task = tk.task_of_group[grp];
    const uint16_t* apT = ap + (int64_t)task * VP * C;                     // [VP][C]
    const uint16_t* bmT = bmt + (int64_t)task * C * VP;                    // [C][VP]
    const float* caT = ca + task * VP;
    // this lane's rows (clamped: rows past the group's end are computed on the last row and not stored)
    const bool live_a = local0 + g < rows_per_group, live_b = R16 && local0 + g + 8 < rows_per_group;
    const int64_t last = (int64_t)grp * rows_per_group + rows_per_group - 1;
    const int64_t row_a = live_a ? (int64_t)grp * rows_per_group + local0 + g : last;
    const int64_t row_b = live_b ? (int64_t)grp * rows_per_group + local0 + g + 8 : last;
    // The contraction index of an MMA may be permuted freely as long as both operands agree, and so may the columns of
    // its output: within every 16-wide block, fragment positions (2tq, 2tq+1, 2tq+8, 2tq+9) are mapped to the FOUR
    // CONSECUTIVE columns 4tq .. 4tq+3, so a lane moves one 16-byte piece per row and block.
    float* xa = hs + row_a * ldh + 4 * tq;
    float* xb = hs + row_b * ldh + 4 * tq;
    constexpr int NB = R16 ? KC : 1;                                       // row b only exists for 16-row warps
    float4 va[KC], vb[NB];
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) va[kk] = ld_stream4(xa + 16 * kk);
    if (R16) {
#pragma unroll
        for (int kk = 0; kk < NB; ++kk) vb[kk] = ld_stream4(xb + 16 * kk);
    }

    // ---- LayerNorm 2 statistics: the four lanes of a quad hold one row; mean first, then the centred second moment
    float mean_a, rstd_a, mean_b = 0.f, rstd_b = 0.f;
    {
        float s = 0.f;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) s += (va[kk].x + va[kk].y) + (va[kk].z + va[kk].w);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        mean_a = s * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float d0 = va[kk].x - mean_a, d1 = va[kk].y - mean_a, d2 = va[kk].z - mean_a, d3 = va[kk].w - mean_a;
            q = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, q))));
        }
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        rstd_a = rsqrtf(q * (1.0f / C) + eps2);
    }
    if (R16) {
        float s = 0.f;
#pragma unroll
        for (int kk = 0; kk < NB; ++kk) s += (vb[kk].x + vb[kk].y) + (vb[kk].z + vb[kk].w);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        mean_b = s * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int kk = 0; kk < NB; ++kk) {
            const float d0 = vb[kk].x - mean_b, d1 = vb[kk].y - mean_b, d2 = vb[kk].z - mean_b, d3 = vb[kk].w - mean_b;
            q = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, q))));
        }
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        rstd_b = rsqrtf(q * (1.0f / C) + eps2);
    }

    // ---- scores[R x VP] = xn[R x C] . ap^T
    float sc[NTILE][4];
#pragma unroll
    for (int t = 0; t < NTILE; ++t) { sc[t][0] = sc[t][1] = sc[t][2] = sc[t][3] = 0.f; }
    const float na = -mean_a * rstd_a, nb_ = -mean_b * rstd_b;
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
        uint32_t af[4];
        af[0] = pack16x2(fmaf(va[kk].x, rstd_a, na), fmaf(va[kk].y, rstd_a, na), FMT);
        af[2] = pack16x2(fmaf(va[kk].z, rstd_a, na), fmaf(va[kk].w, rstd_a, na), FMT);
        if (R16) {
            const float4 b = vb[R16 ? kk : 0];
            af[1] = pack16x2(fmaf(b.x, rstd_b, nb_), fmaf(b.y, rstd_b, nb_), FMT);
            af[3] = pack16x2(fmaf(b.z, rstd_b, nb_), fmaf(b.w, rstd_b, nb_), FMT);
        } else {
            af[1] = af[3] = 0u;
        }
#pragma unroll
        for (int t = 0; t < NTILE; ++t) {
            const uint2 w = __ldg(reinterpret_cast<const uint2*>(apT + (int64_t)(t * 8 + g) * C + 16 * kk + 4 * tq));
            mma_16816(sc[t], af, w.x, w.y, FMT);
        }
    }
    // ---- softmax over the NT tokens of every head; the probabilities become the A fragments of the second GEMM
    uint32_t pf[KV][4];
#pragma unroll
    for (int t = 0; t < NTILE; ++t) {
        const float2 cc = __ldg(reinterpret_cast<const float2*>(caT + t * 8 + 2 * tq));      // -inf on padding vectors
        float v0 = sc[t][0] + cc.x, v1 = sc[t][1] + cc.y, v2 = sc[t][2] + cc.x, v3 = sc[t][3] + cc.y;
        float ma = fmaxf(v0, v1), mb = fmaxf(v2, v3);
#pragma unroll
        for (int o = 1; o < NT / 2; o <<= 1) {
            ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, o));
            mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, o));
        }
        ma = (ma == -INFINITY) ? 0.f : ma;                                   // a head made of padding only
        mb = (mb == -INFINITY) ? 0.f : mb;
        v0 = __expf(v0 - ma); v1 = __expf(v1 - ma); v2 = __expf(v2 - mb); v3 = __expf(v3 - mb);
        float la = v0 + v1, lb = v2 + v3;
#pragma unroll
        for (int o = 1; o < NT / 2; o <<= 1) {
            la += __shfl_xor_sync(0xffffffffu, la, o);
            lb += __shfl_xor_sync(0xffffffffu, lb, o);
        }
        const float ia = la > 0.f ? 1.0f / la : 0.f, ib = lb > 0.f ? 1.0f / lb : 0.f;
        pf[t >> 1][(t & 1) * 2] = pack16x2(v0 * ia, v1 * ia, FMT);
        pf[t >> 1][(t & 1) * 2 + 1] = R16 ? pack16x2(v2 * ib, v3 * ib, FMT) : 0u;
    }
    // ---- h += bo + P . Bm, one 16-column block (two MMA column tiles) at a time, in registers.
    // Column tile 0 of a block produces channels 4q, 4q+1 and tile 1 channels 4q+2, 4q+3 of every quad q (see above); the
    // vector axis of bmt is stored in the matching fragment order by ops.xattn_tables.
    const int chq = 4 * (g >> 1) + (g & 1);                               // channel (within the block) of B column g, tile 0
#pragma unroll
    for (int nb = 0; nb < KC; ++nb) {
        float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
        const uint16_t* bp0 = bmT + (int64_t)(16 * nb + chq) * VP + 4 * tq;
#pragma unroll
        for (int kv = 0; kv < KV; ++kv) {
            const uint2 w0 = __ldg(reinterpret_cast<const uint2*>(bp0 + 16 * kv));
            const uint2 w1 = __ldg(reinterpret_cast<const uint2*>(bp0 + 2 * VP + 16 * kv));
            mma_16816(d0, pf[kv], w0.x, w0.y, FMT);
            mma_16816(d1, pf[kv], w1.x, w1.y, FMT);
        }
        const float4 bias = __ldg(reinterpret_cast<const float4*>(bo + 16 * nb + 4 * tq));
        va[nb].x += bias.x + d0[0]; va[nb].y += bias.y + d0[1]; va[nb].z += bias.z + d1[0]; va[nb].w += bias.w + d1[1];
        if (live_a) *reinterpret_cast<float4*>(xa + 16 * nb) = va[nb];
        if (R16) {
            float4& b = vb[R16 ? nb : 0];
            b.x += bias.x + d0[2]; b.y += bias.y + d0[3]; b.z += bias.z + d1[2]; b.w += bias.w + d1[3];
            if (live_b) *reinterpret_cast<float4*>(xb + 16 * nb) = b;
        }
    }
    // ---- LayerNorm 3 of the updated rows -> 16-bit operand of the feed-forward
    float m3a, r3a, m3b = 0.f, r3b = 0.f;
    {
        float s = 0.f;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) s += (va[kk].x + va[kk].y) + (va[kk].z + va[kk].w);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        m3a = s * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            const float d0 = va[kk].x - m3a, d1 = va[kk].y - m3a, d2 = va[kk].z - m3a, d3 = va[kk].w - m3a;
            q = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, q))));
        }
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        r3a = rsqrtf(q * (1.0f / C) + eps3);
    }
    if (R16) {
        float s = 0.f;
#pragma unroll
        for (int kk = 0; kk < NB; ++kk) s += (vb[kk].x + vb[kk].y) + (vb[kk].z + vb[kk].w);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        m3b = s * (1.0f / C);
        float q = 0.f;
#pragma unroll
        for (int kk = 0; kk < NB; ++kk) {
            const float d0 = vb[kk].x - m3b, d1 = vb[kk].y - m3b, d2 = vb[kk].z - m3b, d3 = vb[kk].w - m3b;
            q = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, q))));
        }
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        r3b = rsqrtf(q * (1.0f / C) + eps3);
    }
    uint16_t* oa = out + row_a * ldo + 4 * tq;
    uint16_t* ob = out + row_b * ldo + 4 * tq;
#pragma unroll
    for (int nb = 0; nb < KC; ++nb) {
        const float4 gg = __ldg(reinterpret_cast<const float4*>(g3 + 16 * nb + 4 * tq)), bb = __ldg(reinterpret_cast<const float4*>(b3 + 16 * nb + 4 * tq));
        if (live_a) {
            const float4 a = va[nb];
            *reinterpret_cast<uint2*>(oa + 16 * nb) =
                make_uint2(pack16x2((a.x - m3a) * r3a * gg.x + bb.x, (a.y - m3a) * r3a * gg.y + bb.y, FMT),
                           pack16x2((a.z - m3a) * r3a * gg.z + bb.z, (a.w - m3a) * r3a * gg.w + bb.w, FMT));
        }
        if (R16 && live_b) {
            const float4 b = vb[R16 ? nb : 0];
            *reinterpret_cast<uint2*>(ob + 16 * nb) =
                make_uint2(pack16x2((b.x - m3b) * r3b * gg.x + bb.x, (b.y - m3b) * r3b * gg.y + bb.y, FMT),
                           pack16x2((b.z - m3b) * r3b * gg.z + bb.z, (b.w - m3b) * r3b * gg.w + bb.w, FMT));
        }
    }
}

struct TaskIds {
    int main_task[SMTL_MAX_TASKS];
    int src_task[SMTL_MAX_TASKS];
};
// per-pixel cross-task attention: one thread per (q row, head); Nk <= 8 source streams
__global__ void taskattn_kernel(const uint16_t* __restrict__ q, const uint16_t* __restrict__ k,
                                const uint16_t* __restrict__ v, uint16_t* __restrict__ out, int c, int nheads,
                                int n_main, int n_src, int64_t rows_per_group, TaskIds ids, int exclude_self,
                                float scale, int fmt) {
    const int64_t total = (int64_t)n_main * rows_per_group * nheads;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    // main task fastest: the n_main threads of one (pixel, head) sit next to each other, so their reads of the same
    // K/V rows coalesce into one request instead of n_main separate trips to L2/HBM
    const int grp = (int)(idx % n_main);
    const int64_t t2 = idx / n_main;
    const int hd = (int)(t2 % nheads);
    const int64_t pix = t2 / nheads;
    const int64_t row = (int64_t)grp * rows_per_group + pix;
    const int dh = c / nheads;
    const int nch = dh >> 3;
    const int my_task = ids.main_task[grp];
    const uint4* qp = reinterpret_cast<const uint4*>(q + row * c + hd * dh);
    float sc[SMTL_MAX_TASKS];
    float mx = -INFINITY;
#pragma unroll
    for (int s = 0; s < SMTL_MAX_TASKS; ++s) {
        sc[s] = -INFINITY;
        if (s < n_src && ids.src_task[s] >= 0 && !(exclude_self && ids.src_task[s] == my_task)) {   // < 0: empty slot
            const uint4* kp = reinterpret_cast<const uint4*>(k + ((int64_t)s * rows_per_group + pix) * c + hd * dh);
            float d = 0.f;
            for (int i = 0; i < nch; ++i) {
                const uint4 a = __ldg(qp + i), b = __ldg(kp + i);
                float2 x, y;
                x = unpack16x2(a.x, fmt); y = unpack16x2(b.x, fmt); d += x.x * y.x + x.y * y.y;
                x = unpack16x2(a.y, fmt); y = unpack16x2(b.y, fmt); d += x.x * y.x + x.y * y.y;
                x = unpack16x2(a.z, fmt); y = unpack16x2(b.z, fmt); d += x.x * y.x + x.y * y.y;
                x = unpack16x2(a.w, fmt); y = unpack16x2(b.w, fmt); d += x.x * y.x + x.y * y.y;
            }
            sc[s] = d * scale;
            mx = fmaxf(mx, sc[s]);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int s = 0; s < SMTL_MAX_TASKS; ++s) {
        sc[s] = (sc[s] == -INFINITY) ? 0.f : __expf(sc[s] - mx);
        sum += sc[s];
    }
    const float inv = (sum > 0.f) ? 1.0f / sum : 0.f;
    uint4* op = reinterpret_cast<uint4*>(out + row * c + hd * dh);
    for (int i = 0; i < nch; ++i) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int s = 0; s < SMTL_MAX_TASKS; ++s) {
            if (s < n_src && sc[s] != 0.f) {
                const uint4 b =
                    __ldg(reinterpret_cast<const uint4*>(v + ((int64_t)s * rows_per_group + pix) * c + hd * dh) + i);
                float2 y;
                y = unpack16x2(b.x, fmt); acc[0] += sc[s] * y.x; acc[1] += sc[s] * y.y;
                y = unpack16x2(b.y, fmt); acc[2] += sc[s] * y.x; acc[3] += sc[s] * y.y;
                y = unpack16x2(b.z, fmt); acc[4] += sc[s] * y.x; acc[5] += sc[s] * y.y;
                y = unpack16x2(b.w, fmt); acc[6] += sc[s] * y.x; acc[7] += sc[s] * y.y;
            }
        }
        uint4 o;
        o.x = pack16x2(acc[0] * inv, acc[1] * inv, fmt); o.y = pack16x2(acc[2] * inv, acc[3] * inv, fmt);
        o.z = pack16x2(acc[4] * inv, acc[5] * inv, fmt); o.w = pack16x2(acc[6] * inv, acc[7] * inv, fmt);
        op[i] = o;
    }
}

// ============================================================================================= evaluation pre-reductions
// least-squares alignment sums: grid (blocks per image, batch); fp64 accumulation, one atomicAdd(double) x 5 per warp
__global__ void __launch_bounds__(256) lsqsums_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                      const uint8_t* __restrict__ valid, int64_t hw,
                                                      double* __restrict__ sums) {
    const int b = blockIdx.y;
    const float* p = pred + (int64_t)b * hw;
    const float* g = gt + (int64_t)b * hw;
    const uint8_t* v = valid ? valid + (int64_t)b * hw : nullptr;
    double n = 0.0, sp = 0.0, sg = 0.0, spp = 0.0, spg = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
        if (v && !v[i]) continue;
        const double x = (double)__ldg(p + i), y = (double)__ldg(g + i);
        n += 1.0; sp += x; sg += y; spp += x * x; spg += x * y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        sp += __shfl_xor_sync(0xffffffffu, sp, o);
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
        spp += __shfl_xor_sync(0xffffffffu, spp, o);
        spg += __shfl_xor_sync(0xffffffffu, spg, o);
    }
    if ((threadIdx.x & 31) == 0) {
        double* dst = sums + (int64_t)b * 5;
        atomicAdd(dst, n); atomicAdd(dst + 1, sp); atomicAdd(dst + 2, sg); atomicAdd(dst + 3, spp); atomicAdd(dst + 4, spg);
    }
}

// confusion histogram: per-CTA shared-memory bins (integer atomics: the result is order-independent, bit-exact)
__global__ void __launch_bounds__(256) confusion_kernel(const int64_t* __restrict__ lt, const int64_t* __restrict__ lp,
                                                        const uint8_t* __restrict__ valid, int64_t n, int nc,
                                                        unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned int bins[];          // nc * nc + 1
    const int nb = nc * nc + 1;
    for (int i = threadIdx.x; i < nb; i += blockDim.x) bins[i] = 0u;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (valid && !valid[i]) continue;
        const int64_t t = __ldg(lt + i), p = __ldg(lp + i);
        if (t < 0 || t >= nc) continue;
        if (p < 0 || p >= nc) { atomicAdd(&bins[nc * nc], 1u); continue; }
        atomicAdd(&bins[(int)t * nc + (int)p], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += blockDim.x)
        if (bins[i]) atomicAdd(hist + i, (unsigned long long)bins[i]);
}

__global__ void chanmix_kernel(const float* __restrict__ x, int64_t rows, int cin, int cout, const float* __restrict__ w,
                               const float* __restrict__ b, float* __restrict__ y) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float in[16];
        for (int i = 0; i < cin; ++i) in[i] = x[r * cin + i];
        for (int j = 0; j < cout; ++j) {
            float acc = b ? b[j] : 0.f;
            for (int i = 0; i < cin; ++i) acc += in[i] * w[j * cin + i];
            y[r * cout + j] = acc;
        }
    }
}

// ============================================================================================= narrow-output conv head
// A 3x3 conv with 3-4 output channels (the VAE decoder's conv_out, 128 -> 3) as a 9-tap implicit GEMM reads every
// activation from shared memory once per tap for a handful of output columns: 26 TFLOP/s.  Instead the three kx taps are
// folded into the GEMM's N side -- P[row, kx * cout + co] = sum_ky sum_c a[row + (ky - 1) * (w + 2), c] W[co, c, ky, kx],
// a 3-segment implicit GEMM with N = 3 * cout -- and this kernel adds the three horizontally shifted entries:
// out[pixel, co] = bias[co] + sum_kx P[pixel + kx - 1, kx * cout + co]  (rows of P are 16 floats: the three reads of a
// pixel and of its neighbours are contiguous).
__global__ void __launch_bounds__(256) conv_head_gather_kernel(const float* __restrict__ part, int ldp, int batch, int h, int w,
                                                               int cout, const float* __restrict__ bias,
                                                               float* __restrict__ out) {
    const int wp = w + 2;
    const int64_t hw = (int64_t)h * w, plane = (int64_t)(h + 2) * wp;
    const int64_t total = (int64_t)batch * hw;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = idx / hw;
        const int p = (int)(idx - b * hw);
        const int y = p / w, x = p - y * w;
        float acc[4];
#pragma unroll
        for (int co = 0; co < 4; ++co) acc[co] = (co < cout && bias) ? __ldg(bias + co) : 0.f;
        const float* base = part + (b * plane + (int64_t)(y + 1) * wp + x) * ldp;   // padded pixel left of (y, x): kx = 0
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float* src = base + (int64_t)kx * ldp + kx * cout;
#pragma unroll
            for (int co = 0; co < 4; ++co)
                if (co < cout) acc[co] += __ldg(src + co);
        }
        for (int co = 0; co < cout; ++co) out[idx * cout + co] = acc[co];
    }
}

// ============================================================================================= task-map epilogue
__device__ __forceinline__ float clip1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }

__global__ void taskmap_kernel(const float* __restrict__ x, int batch, int hw, int mode, float* __restrict__ out_clip,
                               float* __restrict__ out_post, long long* __restrict__ out_ids,
                               const float* __restrict__ palette, int npal) {
    const int64_t total = (int64_t)batch * hw;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = idx / hw, p = idx - b * hw;
        const float c0 = x[idx * 3], c1 = x[idx * 3 + 1], c2 = x[idx * 3 + 2];
        if (mode == SMTL_MAP_MEAN1) {
            // torch mean over 3 channels = (c0 + c1 + c2) / 3 (stablemtl_pipeline.py:647), then clip (:601)
            const float m = clip1((c0 + c1 + c2) / 3.0f);
            if (out_clip) out_clip[idx] = m;
            if (out_post) out_post[idx] = (m + 1.0f) / 2.0f;
        } else if (mode == SMTL_MAP_FLOW2) {
            const float a = clip1(c0), bb = clip1(c1);
            if (out_clip) { out_clip[(b * 2) * hw + p] = a; out_clip[(b * 2 + 1) * hw + p] = bb; }
            if (out_post) { out_post[(b * 2) * hw + p] = a; out_post[(b * 2 + 1) * hw + p] = bb; }
        } else {
            const float a = clip1(c0), bb = clip1(c1), cc = clip1(c2);
            if (out_clip) {
                out_clip[(b * 3) * hw + p] = a; out_clip[(b * 3 + 1) * hw + p] = bb; out_clip[(b * 3 + 2) * hw + p] = cc;
            }
            if (mode == SMTL_MAP_SEMANTIC) {
                if (out_ids) {
                    int best = 0;
                    float bd = INFINITY;
                    for (int k = 0; k < npal; ++k) {
                        const float d0 = a - palette[3 * k], d1 = bb - palette[3 * k + 1], d2 = cc - palette[3 * k + 2];
                        const float d = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);   // torch.cdist p=2, argmin first index
                        if (d < bd) { bd = d; best = k; }
                    }
                    out_ids[idx] = best;
                }
            } else if (out_post) {
                float r0 = a, r1 = bb, r2 = cc;
                if (mode == SMTL_MAP_RGB3) {
                    r0 = (a + 1.0f) / 2.0f; r1 = (bb + 1.0f) / 2.0f; r2 = (cc + 1.0f) / 2.0f;
                } else if (mode == SMTL_MAP_NORMAL) {
                    float nrm = sqrtf(a * a + bb * bb + cc * cc);
                    if (nrm == 0.f) nrm = 1.0f;
                    r0 = a / nrm; r1 = bb / nrm; r2 = cc / nrm;
                }
                out_post[(b * 3) * hw + p] = r0; out_post[(b * 3 + 1) * hw + p] = r1; out_post[(b * 3 + 2) * hw + p] = r2;
            }
        }
    }
}

inline int grid_for(int64_t total, int block, int max_blocks = 148 * 16) {
    int64_t g = (total + block - 1) / block;
    if (g > max_blocks) g = max_blocks;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace

// ================================================================================================ C ABI
extern "C" int smtl_gnapply_run(const smtl_gnapply_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->x0 && a->stats0 && a->gamma && a->beta && a->out_bf16, "gnapply: NULL argument");
    const int C = a->c0 + a->c1;
    SMTL_CHECK_ARG(a->c1 == 0 || (a->x1 && a->stats1), "gnapply: c1 > 0 without x1/stats1");
    SMTL_CHECK_ARG(a->groups > 0 && C % a->groups == 0 && a->c0 % 8 == 0 && a->c1 % 8 == 0,
                   "gnapply: C=%d+%d groups=%d unsupported", a->c0, a->c1, a->groups);
    SMTL_CHECK_ARG(C <= 4096 && a->groups <= 64, "gnapply: C=%d too wide", C);
    SMTL_CHECK_ARG(a->batch >= 1 && a->h >= 1 && a->w >= 1 && a->stats_replicas >= 1, "gnapply: bad extent");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int hp = a->pad_out ? a->h + 2 : a->h, wp = a->pad_out ? a->w + 2 : a->w;
    const int cv8 = C / 8;
    SMTL_CHECK_ARG(cv8 <= 320, "gnapply: C=%d too wide", C);
    SMTL_CHECK_ARG((int64_t)hp * wp < ((int64_t)1 << 31), "gnapply: image too large");
    const int threads = (cv8 > 256) ? cv8 : (256 / cv8) * cv8;        // a multiple of C/8: thread <-> fixed channels (>= 129)
    const int ppb = threads / cv8;
    const int64_t steps = ((int64_t)hp * wp + (int64_t)ppb * 4 - 1) / ((int64_t)ppb * 4);
    int bpi = (int)steps;
    const int cap = (148 * 16 + a->batch - 1) / a->batch;
    if (bpi > cap) bpi = cap;
    if (bpi < 1) bpi = 1;
    const size_t smem = (2 * C + 2 * a->groups) * sizeof(float);
    const bool same = (a->pad_out != 0) == (a->x_padded != 0);
    const int variant = (a->x_fmt16 ? 8 : 0) | (a->fmt16 == SMTL_FMT_F16 ? 4 : 0) | (a->raw_bf16 ? 2 : 0) | (same ? 1 : 0);
    using Kern = void (*)(const void*, const void*, int, int, const long long*, const long long*, int, int, int, int, int,
                          float, const float*, const float*, int, int, uint16_t*, uint16_t*, FastDiv, int);
    static const Kern table[16] = {
        gn_apply2_kernel<false, FMT_BF16, false, false>, gn_apply2_kernel<false, FMT_BF16, false, true>,
        gn_apply2_kernel<false, FMT_BF16, true, false>,  gn_apply2_kernel<false, FMT_BF16, true, true>,
        gn_apply2_kernel<false, FMT_F16, false, false>,  gn_apply2_kernel<false, FMT_F16, false, true>,
        gn_apply2_kernel<false, FMT_F16, true, false>,   gn_apply2_kernel<false, FMT_F16, true, true>,
        gn_apply2_kernel<true, FMT_BF16, false, false>,  gn_apply2_kernel<true, FMT_BF16, false, true>,
        gn_apply2_kernel<true, FMT_BF16, true, false>,   gn_apply2_kernel<true, FMT_BF16, true, true>,
        gn_apply2_kernel<true, FMT_F16, false, false>,   gn_apply2_kernel<true, FMT_F16, false, true>,
        gn_apply2_kernel<true, FMT_F16, true, false>,    gn_apply2_kernel<true, FMT_F16, true, true>};
    table[variant]<<<dim3(bpi, a->batch), threads, smem, st>>>(
        a->x0, a->x1, a->c0, a->c1, reinterpret_cast<const long long*>(a->stats0),
        reinterpret_cast<const long long*>(a->stats1), a->stats_replicas, a->batch, a->h, a->w,
        a->groups, a->eps, a->gamma, a->beta, a->silu, a->pad_out, reinterpret_cast<uint16_t*>(a->out_bf16),
        reinterpret_cast<uint16_t*>(a->raw_bf16), make_fastdiv((uint32_t)wp), a->x_padded);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_gnfinalize_run(const smtl_gnfinalize_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->stats && a->gamma && a->beta && a->ss, "gnfinalize: NULL argument");
    SMTL_CHECK_ARG(a->groups > 0 && a->c % a->groups == 0 && a->batch >= 1 && a->stats_replicas >= 1 && a->pixels > 0 &&
                       a->c <= 8192 && a->groups <= 64, "gnfinalize: bad extent");
    const size_t smem = (2 * a->groups) * sizeof(float);
    const int threads = a->c < 256 ? ((a->c + 31) / 32) * 32 : 256;
    gn_finalize_kernel<<<a->batch, threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const long long*>(a->stats), a->stats_replicas, a->batch, a->c, a->groups, (double)a->pixels * (a->c / a->groups), a->eps, a->gamma,
        a->beta, a->ss);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_lsqsums_run(const smtl_lsqsums_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->pred && a->gt && a->sums, "lsqsums: NULL argument");
    SMTL_CHECK_ARG(a->batch > 0 && a->batch <= 65535 && a->hw > 0, "lsqsums: bad extent");
    int64_t blocks = (a->hw + 256 * 8 - 1) / (256 * 8);
    if (blocks > 1184) blocks = 1184;               // 8 CTAs per SM
    lsqsums_kernel<<<dim3((unsigned)blocks, (unsigned)a->batch), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        a->pred, a->gt, a->valid, a->hw, a->sums);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_confusion_run(const smtl_confusion_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->label_true && a->label_pred && a->hist, "confusion: NULL argument");
    SMTL_CHECK_ARG(a->n > 0 && a->n_classes > 0 && a->n_classes <= 64, "confusion: bad extent (n_classes <= 64)");
    int64_t blocks = (a->n + 256 * 16 - 1) / (256 * 16);
    if (blocks > 1184) blocks = 1184;
    const size_t smem = ((size_t)a->n_classes * a->n_classes + 1) * sizeof(unsigned int);
    confusion_kernel<<<(unsigned)blocks, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        a->label_true, a->label_pred, a->valid, a->n, a->n_classes, reinterpret_cast<unsigned long long*>(a->hist));
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_memset_run(const smtl_memset_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->ptr && a->bytes > 0, "memset: NULL / empty");
    SMTL_CHECK_CUDA(cudaMemsetAsync(a->ptr, a->value, (size_t)a->bytes, reinterpret_cast<cudaStream_t>(stream)));
    return SMTL_OK;
}

extern "C" int smtl_ln_run(const smtl_ln_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->x && a->gamma0 && a->beta0 && a->out0, "ln: NULL argument");
    SMTL_CHECK_ARG(a->c % 4 == 0 && a->c <= 1280 && a->ldx % 4 == 0 && a->ldo % 4 == 0, "ln: c=%d ldx=%d unsupported",
                   a->c, a->ldx);
    SMTL_CHECK_ARG(a->rows > 0 && a->rows_per_group > 0, "ln: bad rows");
    SMTL_CHECK_ARG(!a->out1 || (a->gamma1 && a->beta1), "ln: out1 without affine");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (a->x_is_bf16) {
        if (a->c <= 384) launch_ln<true, 3, 4>(a, st);
        else if (a->c <= 640) launch_ln<true, 5, 2>(a, st);
        else launch_ln<true, 10, 1>(a, st);
    } else {
        if (a->c <= 384) launch_ln<false, 3, 4>(a, st);
        else if (a->c <= 640) launch_ln<false, 5, 2>(a, st);
        else launch_ln<false, 10, 1>(a, st);
    }
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_upsample_run(const smtl_upsample_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->x && a->out_bf16, "upsample: NULL argument");
    SMTL_CHECK_ARG(a->c % 8 == 0 && a->oh >= a->h && a->ow >= a->w, "upsample: bad shape");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t total = (int64_t)a->batch * (a->oh + 2) * (a->ow + 2) * (a->c / 8);
    upsample_pad_kernel<<<grid_for(total, 256), 256, 0, st>>>(a->x, a->x_fmt16, a->batch, a->h, a->w, a->c, a->oh,
                                                              a->ow, (uint16_t*)a->out_bf16, a->fmt16);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_im2col_run(const smtl_im2col_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->x && a->out_bf16, "im2col: NULL argument");
    SMTL_CHECK_ARG(a->kpad >= 9 * a->c && a->kpad % 8 == 0, "im2col: kpad %d < 9*c or unaligned", a->kpad);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (a->c % 8 == 0 && a->kpad == 9 * a->c) {
        const int64_t total = (int64_t)a->batch * a->oh * a->ow * 9 * (a->c / 8);
        im2col_vec8_kernel<<<grid_for(total, 256), 256, 0, st>>>(a->x, a->x_fmt16, a->batch, a->h, a->w, a->c, a->stride,
                                                                 a->pad_t, a->pad_l, a->oh, a->ow,
                                                                 (uint16_t*)a->out_bf16, a->fmt16);
    } else {
        SMTL_CHECK_ARG(!a->x_fmt16, "im2col: the scalar (tiny Cin) path takes fp32 input");
        const int64_t total = (int64_t)a->batch * a->oh * a->ow * (a->kpad / 8);
        im2col_scalar_kernel<<<grid_for(total, 256), 256, 0, st>>>((const float*)a->x, a->batch, a->h, a->w, a->c, a->stride,
                                                                   a->pad_t, a->pad_l, a->oh, a->ow, a->kpad,
                                                                   (uint16_t*)a->out_bf16, a->fmt16);
    }
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_rgbprep_run(const smtl_rgbprep_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->rgb_nchw && a->out_nhwc, "rgbprep: NULL argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t total = (int64_t)a->batch * a->h * a->w;
    rgbprep_kernel<<<grid_for(total, 256), 256, 0, st>>>(a->rgb_nchw, a->src_u8, a->batch, a->h * a->w, a->out_nhwc);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_rgbstem_run(const smtl_rgbstem_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->rgb_nchw && a->out_bf16, "rgbstem: NULL argument");
    SMTL_CHECK_ARG(a->batch > 0 && a->h > 0 && a->w > 0 && a->src_mode >= 0 && a->src_mode <= 2, "rgbstem: bad extent / mode");
    const int64_t total = (int64_t)a->batch * a->h * a->w;
    rgb_stem_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        a->rgb_nchw, a->src_mode, a->batch, a->h, a->w, reinterpret_cast<uint16_t*>(a->out_bf16), a->fmt16);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_unetin_run(const smtl_unetin_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->latents && a->first_img && a->second_img && a->out, "unetin: NULL argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t total = (int64_t)a->out_images * a->hw;
    unetin_kernel<<<grid_for(total, 256), 256, 0, st>>>(a->latents, a->first_img, a->second_img, a->out_images, a->hw,
                                                        a->out);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_softmax_run(const smtl_softmax_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->s && a->p_bf16, "softmax: NULL argument");
    SMTL_CHECK_ARG(a->n > 0 && a->n <= 8192 && a->rows > 0, "softmax: n=%d out of range", a->n);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool vec = a->n % 4 == 0 && a->lds % 4 == 0 && a->ldp % 4 == 0 && (reinterpret_cast<uintptr_t>(a->s) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(a->p_bf16) & 7) == 0;
    if (vec)
        softmax_rows_vec4_kernel<<<(unsigned)a->rows, 128, 0, st>>>(a->s, a->n, a->lds, a->scale, (uint16_t*)a->p_bf16,
                                                                    a->ldp, a->fmt16);
    else
        softmax_rows_kernel<<<(unsigned)a->rows, 256, 0, st>>>(a->s, a->n, a->lds, a->scale, (uint16_t*)a->p_bf16,
                                                               a->ldp, a->fmt16);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_xattn_run(const smtl_xattn_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->q_bf16 && a->kc && a->vc && a->out_bf16, "xattn: NULL argument");
    SMTL_CHECK_ARG(a->rows > 0 && a->rows_per_group > 0 && a->heads > 0, "xattn: bad extent");
    SMTL_CHECK_ARG(a->ldq % 8 == 0 && a->ldo % 8 == 0, "xattn: row strides must be multiples of 8 elements");
    SMTL_CHECK_ARG((a->rows + a->rows_per_group - 1) / a->rows_per_group <= SMTL_MAX_TASKS, "xattn: too many groups");
    SMTL_CHECK_ARG(a->ntok_pad >= 1 && a->ntok_pad <= SMTL_MAX_XATTN_TOKENS, "xattn: ntok_pad %d (1..%d)", a->ntok_pad,
                   SMTL_MAX_XATTN_TOKENS);
    XattnK tk;
    for (int i = 0; i < SMTL_MAX_TASKS; ++i) {
        tk.ntok[i] = a->ntok[i];
        tk.task_of_group[i] = a->task_of_group[i];
        SMTL_CHECK_ARG(a->ntok[i] >= 0 && a->ntok[i] <= a->ntok_pad, "xattn: ntok[%d]=%d", i, a->ntok[i]);
        SMTL_CHECK_ARG(a->task_of_group[i] >= 0 && a->task_of_group[i] < SMTL_MAX_TASKS, "xattn: bad task id");
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int wpb = 8;
    const int64_t grid = (a->rows + wpb - 1) / wpb;
    xattn_kernel<<<(unsigned)grid, wpb * 32, 0, st>>>((const uint16_t*)a->q_bf16, a->ldq, a->rows, a->heads, a->kc,
                                                      a->vc, tk, a->ntok_pad, a->rows_per_group, (uint16_t*)a->out_bf16, a->ldo,
                                                      a->scale, a->fmt16);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

template <int H, int NT>
static int launch_xattn_mma(const smtl_xattnf_args* a, const XattnK& tk, int ngroups, cudaStream_t st) {
    constexpr bool R16 = H <= 5;                   // 160 fp32 of row data per lane: 16 rows up to 320 channels, 8 rows at 640
    constexpr int R = R16 ? 16 : 8;
    const int64_t wpg = (a->rows_per_group + R - 1) / R;
    const int64_t warps = wpg * ngroups;
    auto kern = a->fmt16 == SMTL_FMT_F16 ? xattn_mma_kernel<H, NT, FMT_F16, R16> : xattn_mma_kernel<H, NT, FMT_BF16, R16>;
    kern<<<(unsigned)((warps + 3) / 4), 128, 0, st>>>(
        a->hs, a->ldh, a->rows_per_group, ngroups, tk, (const uint16_t*)a->ap, a->ca, (const uint16_t*)a->bmt, a->bo,
        a->gamma3, a->beta3, (uint16_t*)a->out_bf16, a->ldo, a->eps2, a->eps3);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_xattnf_supported(int32_t heads, int32_t ntok_pad) {
    return (heads == 5 || heads == 10 || heads == 1 || heads == 2) && (ntok_pad == 4 || ntok_pad == 8);
}

extern "C" int smtl_xattnf_run(const smtl_xattnf_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->hs && a->ap && a->ca && a->bmt && a->bo && a->gamma3 && a->beta3 && a->out_bf16,
                   "xattnf: NULL argument");
    SMTL_CHECK_ARG(a->rows > 0 && a->rows_per_group > 0 && a->rows % a->rows_per_group == 0, "xattnf: bad rows");
    const int ngroups = (int)(a->rows / a->rows_per_group);
    SMTL_CHECK_ARG(ngroups <= SMTL_MAX_TASKS, "xattnf: too many row groups");
    SMTL_CHECK_ARG(a->ldh % 4 == 0 && a->ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(a->hs) & 15) == 0 &&
                       (reinterpret_cast<uintptr_t>(a->out_bf16) & 7) == 0,
                   "xattnf: rows must be 16-byte (hs) / 8-byte (out) aligned");
    SMTL_CHECK_ARG(smtl_xattnf_supported(a->heads, a->ntok_pad), "xattnf: heads=%d ntok_pad=%d not instantiated", a->heads,
                   a->ntok_pad);
    XattnK tk;
    for (int i = 0; i < SMTL_MAX_TASKS; ++i) {
        tk.ntok[i] = 0;
        tk.task_of_group[i] = a->task_of_group[i];
        SMTL_CHECK_ARG(a->task_of_group[i] >= 0 && a->task_of_group[i] < SMTL_MAX_TASKS, "xattnf: bad task id");
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const bool n8 = a->ntok_pad == 8;
    switch (a->heads) {
        case 1: return n8 ? launch_xattn_mma<1, 8>(a, tk, ngroups, st) : launch_xattn_mma<1, 4>(a, tk, ngroups, st);
        case 2: return n8 ? launch_xattn_mma<2, 8>(a, tk, ngroups, st) : launch_xattn_mma<2, 4>(a, tk, ngroups, st);
        case 5: return n8 ? launch_xattn_mma<5, 8>(a, tk, ngroups, st) : launch_xattn_mma<5, 4>(a, tk, ngroups, st);
        default: return n8 ? launch_xattn_mma<10, 8>(a, tk, ngroups, st) : launch_xattn_mma<10, 4>(a, tk, ngroups, st);
    }
}

extern "C" int smtl_taskattn_run(const smtl_taskattn_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->q_bf16 && a->k_bf16 && a->v_bf16 && a->out_bf16, "taskattn: NULL argument");
    SMTL_CHECK_ARG(a->n_main >= 1 && a->n_main <= SMTL_MAX_TASKS && a->n_src >= 1 && a->n_src <= SMTL_MAX_TASKS,
                   "taskattn: n_main=%d n_src=%d", a->n_main, a->n_src);
    SMTL_CHECK_ARG(a->c % (a->nheads * 8) == 0, "taskattn: c=%d nheads=%d", a->c, a->nheads);
    TaskIds ids;
    for (int i = 0; i < SMTL_MAX_TASKS; ++i) { ids.main_task[i] = a->main_task[i]; ids.src_task[i] = a->src_task[i]; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t total = (int64_t)a->n_main * a->rows_per_group * a->nheads;
    const int64_t grid = (total + 127) / 128;
    taskattn_kernel<<<(unsigned)grid, 128, 0, st>>>((const uint16_t*)a->q_bf16, (const uint16_t*)a->k_bf16,
                                                    (const uint16_t*)a->v_bf16, (uint16_t*)a->out_bf16, a->c,
                                                    a->nheads, a->n_main, a->n_src, a->rows_per_group, ids,
                                                    a->exclude_self, a->scale, a->fmt16);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_chanmix_run(const smtl_chanmix_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->x && a->w && a->y, "chanmix: NULL argument");
    SMTL_CHECK_ARG(a->cin >= 1 && a->cin <= 16 && a->cout >= 1 && a->cout <= 16 && a->rows > 0, "chanmix: bad shape");
    chanmix_kernel<<<grid_for(a->rows, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a->x, a->rows, a->cin,
                                                                                              a->cout, a->w, a->b, a->y);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_headgather_run(const smtl_headgather_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->partial && a->out, "headgather: NULL argument");
    SMTL_CHECK_ARG(a->batch > 0 && a->h > 0 && a->w > 0 && a->cout >= 1 && a->cout <= 4 && a->ldp >= 3 * a->cout,
                   "headgather: bad extent (cout <= 4, ldp >= 3 * cout)");
    const int64_t total = (int64_t)a->batch * a->h * a->w;
    conv_head_gather_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        a->partial, a->ldp, a->batch, a->h, a->w, a->cout, a->bias, a->out);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}

extern "C" int smtl_taskmap_run(const smtl_taskmap_args* a, void* stream) {
    SMTL_CHECK_ARG(a && a->x, "taskmap: NULL argument");
    SMTL_CHECK_ARG(a->mode >= SMTL_MAP_MEAN1 && a->mode <= SMTL_MAP_SEMANTIC, "taskmap: mode %d", a->mode);
    SMTL_CHECK_ARG(a->mode != SMTL_MAP_SEMANTIC || !a->out_ids || (a->palette && a->npalette > 0),
                   "taskmap: semantic needs a palette");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t total = (int64_t)a->batch * a->hw;
    taskmap_kernel<<<grid_for(total, 256), 256, 0, st>>>(a->x, a->batch, a->hw, a->mode, a->out_clipped, a->out_post,
                                                         (long long*)a->out_ids, a->palette, a->npalette);
    SMTL_CHECK_CUDA(cudaGetLastError());
    return SMTL_OK;
}
