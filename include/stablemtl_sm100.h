/*
 * stablemtl_sm100.h -- C ABI of libstablemtl_sm100.so: the B200 (sm_100a) kernels behind the
 * single-step latent pass of StableMTLPipeline (VAE encode -> SD-2 UNet [+ task attention] -> VAE decode).
 *
 * The reference has no FFI of its own (it is pure Python on top of torch/cuDNN/cuBLAS/xformers), so every
 * entry point below names the reference call site (file:line under /root/reference) whose library kernel it
 * replaces.  Conventions:
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless a name ends in _host;
 *   - every function returns 0 on success and a negative SMTL_E* code on failure (no exceptions cross the
 *     ABI); smtl_last_error() returns a thread-local message for the last failure;
 *   - no allocation inside: the caller owns every buffer, scratch included;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous and thread-safe w.r.t. distinct
 *     streams; there is no global mutable state;
 *   - activations are pixel-major ("NHWC"): a feature map is a row-major matrix [B*H*W, C]
 *       compact layout : row = (b*H + y)*W + x                  fp32 residual stream / bf16 tokens
 *       padded  layout : row = (b*(H+2) + y+1)*(W+2) + x+1      bf16, 1-pixel zero halo (conv operand)
 *   - weights are bf16 [N, K] K-major; conv filters are [Cout, tap*Cin + c], tap = ky*3 + kx.
 */
#ifndef STABLEMTL_SM100_H
#define STABLEMTL_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMTL_ABI_VERSION 3

enum {
    SMTL_OK = 0,
    SMTL_EINVAL = -1,   /* bad argument (shape/alignment/NULL) */
    SMTL_ECUDA = -2,    /* CUDA runtime / driver error */
    SMTL_ENODEV = -3,   /* no sm_100 device / driver entry point missing */
    SMTL_EKIND = -4     /* unknown op kind in a plan */
};

enum { SMTL_ACT_NONE = 0, SMTL_ACT_GELU = 1, SMTL_ACT_GEGLU = 2, SMTL_ACT_SILU = 3 };
/* CONV_PAD_UP2: like CONV_PAD, but GEMM row = padded pixel (y, x) of the LOW-resolution map and the output row is
 * pixel (2y + py, 2x + px) of the 2x nearest-upsampled map (one of the four output parities of "upsample then 3x3
 * conv", each of which is a 2x2 conv on the low-resolution input; src/model/resnet.py:58-72, diffusers Upsample2D) */
/* The three *_PAD variants write the OUTPUT (and read the residuals) in the padded layout too, so a chain of convs
 * never leaves it: PAD_KEEP = padded row -> same padded row, halo rows of the output are written as zeros;
 * TO_PAD = GEMM rows are compact pixels (token linears, im2col GEMMs), output row = padded index;
 * UP2_PAD = CONV_PAD_UP2 with a padded (2h+2) x (2w+2) output map.  TO_PAD / UP2_PAD leave the output halo untouched. */
enum { SMTL_ROWMAP_IDENTITY = 0, SMTL_ROWMAP_CONV_PAD = 1, SMTL_ROWMAP_CONV_PAD_UP2 = 2, SMTL_ROWMAP_PAD_KEEP = 3,
       SMTL_ROWMAP_TO_PAD = 4, SMTL_ROWMAP_UP2_PAD = 5 };
/* 16-bit operand format of a call (field `fmt16`): the buffers named *_bf16 hold bf16 (0) or IEEE fp16 (1).
 * Both run at the same tcgen05 kind::f16 rate with fp32 accumulation; fp16 conversions saturate at +-65504. */
enum { SMTL_FMT_BF16 = 0, SMTL_FMT_F16 = 1 };

#define SMTL_MAX_SEG 12
#define SMTL_MAX_TASKS 8
#define SMTL_MAX_XATTN_TOKENS 8   /* tokens of a task prompt ("optical flow" = 4 with BOS/EOS; padded count is 4 or 8) */

/* ------------------------------------------------------------------------------------------------ GEMM / conv
 * D[m, n] = sum over K segments of A_src[m + row_shift, a_col0 + kk] * B[n, kbase + kk]
 * One tcgen05 kernel (TMA-fed, TMEM accumulators) serves
 *   - token linears                    nn.Linear                 src/model/attention.py:185,191,210,212,442-451,460
 *                                      diffusers FeedForward     src/model/attention.py:285,372
 *                                      task MLP / MLPv2          src/model/attention.py:494-495,512,598
 *   - 3x3 stride-1 convs as implicit GEMM over the padded layout (9 row-shifted K segments)
 *                                      InflatedConv3d            src/model/resnet.py:14-16,143,159
 *   - 1x1 shortcut fused as a 10th K segment read from a second source   src/model/resnet.py:172,200
 *   - the VAE convs / attention projections (diffusers AutoencoderKL)    src/stablemtl_pipeline.py:619-620,642-643
 * Epilogue (fused): v = act(acc + bias); aux_bf16 = v; v += res1 + res2; out_f32 = v; out_bf16 = v.
 */
typedef struct smtl_gemm_seg {
    int32_t row_shift;  /* added to the GEMM row to get the A row (may be negative; OOB rows read as 0) */
    int32_t kblocks;    /* number of 64-wide K blocks in this segment */
    int32_t src;        /* 0: a0, 1: a1 */
    int32_t a_col0;     /* first A column of the segment */
} smtl_gemm_seg;

typedef struct smtl_gemm_args {
    const void* a0;     /* bf16 [a0_rows, a0_cols], leading dim a0_ld (elements) */
    const void* a1;     /* optional second A source */
    const void* b;      /* bf16 [n, k], leading dim ldb */
    int64_t a0_rows, a1_rows;
    int32_t a0_cols, a1_cols;
    int32_t a0_ld, a1_ld;
    int64_t m;          /* GEMM rows */
    int32_t n, k, ldb;
    int32_t nseg;       /* 0 => one segment covering k from a0 */
    smtl_gemm_seg seg[SMTL_MAX_SEG];
    const float* bias;  /* fp32 [n] (or [m] if bias_per_row) or NULL */
    int32_t bias_per_row;
    int32_t act;        /* SMTL_ACT_*; GEGLU: B rows are tile-interleaved value|gate, output has n/2 columns */
    const float* res1;  /* fp32 [out_rows, n_out] residuals, leading dim ldres, or NULL */
    const float* res2;
    int32_t ldres;
    float* out_f32;     /* any subset of the three outputs may be NULL */
    void* out_bf16;
    void* aux_bf16;     /* value before the residual add (child-UNet feature tap, attention.py:348-349) */
    int32_t ldc;        /* leading dim of out_f32 / out_bf16 */
    int32_t ld_aux;
    int32_t rowmap;     /* SMTL_ROWMAP_CONV_PAD: GEMM row = padded pixel, output row = compact pixel; halo rows dropped */
    int32_t img_h, img_w; /* interior size for the conv row map (padded size is +2) */
    int32_t block_n;    /* 0 = auto; else 32/64/128/160/192/224/256 */
    int32_t fmt16;      /* SMTL_FMT_* of a0/a1/b/out_bf16/aux_bf16 */
    int32_t res_fmt16;  /* 0: res1/res2 are fp32; 1: they are 16-bit (fmt16), same leading dim ldres */
    /* Fused GroupNorm statistics of the OUTPUT (the input of the next GroupNorm, src/model/resnet.py:177,188):
     * stats is int64 [stats_replicas, stats_images, n_out, 4]: per-(image, channel) sum and sum of squares of the
     * final value (after residual adds) as FIXED-POINT cells (sum_lo, sum_hi, sq_lo, sq_hi; value = lo * 2^-32 +
     * hi * 2^-8), accumulated with integer atomic adds -- exact and order-independent, so the statistics and
     * everything downstream are bit-reproducible.  The caller zeroes it first (smtl_memset_run).
     * image = output row / stats_rows_per_image; m must be stats_images whole images (M tiles restart at every
     * image, which makes an image's statistics independent of its position in the batch).  NULL = off. */
    int32_t stats_replicas;         /* >= 1 copies to spread atomic contention; consumers sum them */
    int64_t* stats;
    int32_t stats_rows_per_image;
    int32_t stats_images;
    int32_t cta_group;  /* 0 = auto; 1 = one CTA per 128-row tile; 2 = CTA pair per 256-row tile (tcgen05 cta_group::2) */
    int32_t up_parity;  /* CONV_PAD_UP2: py * 2 + px */
    /* Grouped GEMM (the per-task MLPs of the task attention, src/util/model.py:102-138): rows
     * [g * group_rows, (g+1) * group_rows) use weight rows [g * n, (g+1) * n) of b (stacked [groups * n, k]) and bias
     * [g * n, (g+1) * n).  group_rows must divide m; when it is not a multiple of the 128-row tile the M tiles restart at
     * every group (like the image-aligned tiles of the statistics producers).  Single CTAs only (cta_group 0 / 1).
     * 0 = off. */
    int64_t group_rows;
    /* Tile order of the persistent CTAs: 0 = auto; 1 = round robin, column tile fastest; 2 = every CTA a contiguous run of
     * the column-tile-major order (auto picks 2 for statistics producers: their per-column sums stay on chip for a whole
     * (image, column tile) run and reach `stats` as one atomic per cell). */
    int32_t tile_order;
    /* Statistics granularity: 0 / 1 = one cell per channel.  2, 4 or 8 = the sums of that many ADJACENT channels land in
     * the cell of the first one and the other cells stay zero -- the same totals for every consumer whose GroupNorm
     * groups are unions of such aligned blocks (a group's cells are summed), with a fraction of the epilogue's
     * cross-lane reduction.  n must be a multiple of it. */
    int32_t stats_group;
} smtl_gemm_args;

typedef struct smtl_gemm_op {
    smtl_gemm_args args;
    uint64_t tmap_a0[16];   /* CUtensorMap images (128 B each) */
    uint64_t tmap_a1[16];
    uint64_t tmap_b[16];
    int32_t block_n;
    int32_t grid;
    int32_t tiles_m, tiles_n;
    int32_t total_kblocks;
    int32_t smem_bytes;
    int32_t cta_group;
    /* shift-grouped mainloop (segments with consecutive row shifts share one activation tile): see smtl_gemm.cu */
    int32_t grouped;
    int32_t ngrp, sp, sw;
    struct { int32_t row_shift, nsub, kblocks, src, a_col0, kb0; } grp[SMTL_MAX_SEG];
    uint64_t tmap_x8[2][16];
    int64_t tile_rpi;       /* image-aligned M tiling: GEMM rows per image (0 = off), M tiles per image */
    int32_t tiles_per_img;
    int32_t pair_split;     /* cta_group 2 over image-aligned tiles of small maps: the pair takes two consecutive 128-row tiles */
} smtl_gemm_op;

int smtl_gemm_plan(const smtl_gemm_args* args, smtl_gemm_op* op);
int smtl_gemm_run(const smtl_gemm_op* op, void* stream);

/* ------------------------------------------------------------------------------------------------ attention
 * Flash-style self-attention, softmax in fp32, tcgen05 for QK^T and PV, S and P kept in tensor memory.
 *   head_dim 64, any number of heads: xformers.ops.memory_efficient_attention at src/model/attention.py:391-397,417;
 *   head_dim 512, one head: the VAE mid-block attention (diffusers Attention(heads=1, dim_head=512) in UNetMidBlock2D,
 *   reached from src/stablemtl_pipeline.py:619-620,642-643).
 * q/k/v live in one 16-bit matrix [batch*ntok, ld] (the fused QKV projection) at column offsets *_col0 + head*head_dim.
 */
typedef struct smtl_fattn_args {
    const void* qkv;
    int32_t ld;
    int32_t q_col0, k_col0, v_col0;
    int32_t batch, ntok, heads;
    void* out_bf16;     /* [batch*ntok, ldo], head h at columns h*64 */
    int32_t ldo;
    float scale;        /* 1/sqrt(head_dim) */
    int32_t fmt16;
    int32_t head_dim;   /* 0 or 64: 64; 512 (heads must be 1) */
} smtl_fattn_args;

typedef struct smtl_fattn_op {
    smtl_fattn_args args;
    uint64_t tmap_qkv[16];  /* CUtensorMap image: [128 x 64] boxes */
    uint64_t tmap_kv[16];   /* head_dim 512: [64 x 64] boxes (each CTA of a pair loads half of every K / V tile) */
    int32_t grid_x, grid_y;
    int32_t smem_bytes;
    int32_t pad_;
} smtl_fattn_op;

int smtl_fattn_plan(const smtl_fattn_args* args, smtl_fattn_op* op);
int smtl_fattn_run(const smtl_fattn_op* op, void* stream);

/* Row softmax fp32 -> bf16 probabilities (VAE mid-block single-head attention, d = 512;
 * diffusers Attention inside AutoencoderKL, reached from src/stablemtl_pipeline.py:619,643). */
typedef struct smtl_softmax_args {
    const float* s;
    int64_t rows;
    int32_t n, lds;
    float scale;
    void* p_bf16;
    int32_t ldp;
    int32_t fmt16;
} smtl_softmax_args;
int smtl_softmax_run(const smtl_softmax_args* a, void* stream);

/* Cross-attention on a few constant text tokens (diffusers Attention as attn2, src/model/attention.py:267-275,360-362).
 * kc/vc are the pre-projected keys/values to_k(text), to_v(text): fp32 [ntask, ntok_pad, heads*64]. */
typedef struct smtl_xattn_args {
    const void* q_bf16;
    int32_t ldq;
    int64_t rows;
    int32_t heads;
    const float* kc;
    const float* vc;
    int32_t ntok[SMTL_MAX_TASKS];
    int32_t task_of_group[SMTL_MAX_TASKS]; /* row group g = row / rows_per_group uses text of task_of_group[g] */
    int64_t rows_per_group;
    void* out_bf16;
    int32_t ldo;
    float scale;
    int32_t fmt16;
    int32_t ntok_pad;       /* rows per task of kc / vc (<= SMTL_MAX_XATTN_TOKENS) */
} smtl_xattn_args;
int smtl_xattn_run(const smtl_xattn_args* a, void* stream);

/* The whole cross-attention step of BasicTransformerBlock (src/model/attention.py:355-373) in one pass over the fp32
 * residual stream:  hs += attn2(LayerNorm2(hs), text[task]);  out = LayerNorm3(hs)  (the 16-bit operand of the
 * feed-forward).  The prompt is constant per task, so attn2 is collapsed at load time into 2 * heads * ntok_pad vectors
 * per task (v = head * ntok_pad + token, padded with zero vectors to VP = the next multiple of 16), LayerNorm2's
 * affine folded in; xn = (hs - mean) * rstd is the normalised row:
 *     ap  [ntask, VP, C] 16-bit   gamma2 * (Wq_head^T k[token, head]) / sqrt(64)
 *     ca  [ntask, VP]    fp32     beta2 . (Wq_head^T k[token, head]) / sqrt(64);  -inf for a padding token / vector
 *     bmt [ntask, C, VP] 16-bit   (Wo[:, head] v[token, head]) transposed
 *     score = xn . ap + ca;  hs += bo + sum_v softmax_token(score)[v] * bmt[:, v]
 * Two skinny GEMMs per 16-row block on mma.sync (fp32 accumulate), four passes over the rows (one from HBM).
 * heads in {1, 2, 5, 10} (C = 64 * heads <= 640), ntok_pad in {4, 8}: smtl_xattnf_supported(). */
typedef struct smtl_xattnf_args {
    float* hs;              /* fp32 [rows, ldh], updated in place */
    int32_t ldh;
    int32_t heads;
    int64_t rows;
    int64_t rows_per_group; /* row group g = row / rows_per_group uses the vectors of task_of_group[g] */
    int32_t task_of_group[SMTL_MAX_TASKS];
    int32_t ntok_pad;
    int32_t fmt16;
    const void* ap;
    const float* ca;
    const void* bmt;
    const float* bo;        /* fp32 [C] attn2.to_out.0.bias */
    const float* gamma3;    /* LayerNorm3 affine, fp32 [C] */
    const float* beta3;
    void* out_bf16;         /* 16-bit [rows, ldo] */
    int32_t ldo;
    float eps2, eps3;
    int32_t pad_;
} smtl_xattnf_args;         /* (no padding: 7 pointers after the two int32 above) */
int smtl_xattnf_run(const smtl_xattnf_args* a, void* stream);
int smtl_xattnf_supported(int32_t heads, int32_t ntok_pad);

/* Per-pixel cross-task attention (src/model/attention.py:500-519,553-597): Nq = 1, Nk = number of other task
 * streams, nheads heads of c/nheads channels. q rows are (main task group, image, pixel); k/v rows are
 * (source task group, image, pixel). */
typedef struct smtl_taskattn_args {
    const void* q_bf16;
    const void* k_bf16;
    const void* v_bf16;
    void* out_bf16;
    int32_t c, nheads;
    int32_t n_main, n_src;
    int64_t rows_per_group;                 /* images * tokens */
    int32_t main_task[SMTL_MAX_TASKS];      /* task id of each q row group */
    int32_t src_task[SMTL_MAX_TASKS];       /* task id of each k/v row group; < 0 = empty slot (rows ignored) */
    int32_t exclude_self;                   /* skip src whose task id equals the row's main task (stablemtl_pipeline.py:483-484) */
    float scale;                            /* 1/sqrt(c/nheads) */
    int32_t fmt16;
    int32_t pad_;
} smtl_taskattn_args;
int smtl_taskattn_run(const smtl_taskattn_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------ normalisation
 * GroupNorm (+SiLU) emitting the 16-bit operand of the next GEMM/conv.  Replaces torch GroupNorm + F.silu + torch.cat at
 * src/model/resnet.py:177-178,188,194, src/model/attention.py:183, src/model/unet.py:438-439,
 * src/model/unet_blocks.py:509,597 and the GroupNorms of diffusers' VAE blocks.
 */
/* GroupNorm apply (+SiLU) from PRODUCER-SIDE statistics: the GEMM/conv that wrote x0 (and x1 of a channel
 * concat, src/model/unet_blocks.py:509,597) also accumulated per-(image, channel) sum / sum of squares
 * (smtl_gemm_args.stats), so this is a single streaming pass: 16-bit (or fp32) compact map in, 16-bit operand of
 * the next conv/GEMM out (zero-halo padded or compact).  Group statistics of a virtual concat are assembled from
 * both sources' channel sums, so a group may straddle the seam. */
typedef struct smtl_gnapply_args {
    const void* x0;
    const void* x1;         /* second source of the concat, or NULL */
    int32_t c0, c1;
    int32_t x_fmt16;        /* 0: x0/x1 are fp32; 1: 16-bit (fmt16) */
    int32_t stats_replicas;
    int32_t x_padded;       /* 1: x0/x1 are in the padded layout [batch, h+2, w+2, c] (halo ignored) */
    int32_t pad2_;
    const int64_t* stats0;  /* int64 fixed-point cells [stats_replicas, batch, c0, 4] (smtl_gemm_args.stats) */
    const int64_t* stats1;  /* [stats_replicas, batch, c1, 4] or NULL */
    int32_t batch, h, w;
    int32_t groups;
    float eps;
    int32_t silu;           /* 0: none; 1: x*sigmoid(x) with sigmoid from one tanh.approx (rel. 2^-11); 2: ex2+rcp (~2 ulp) */
    const float* gamma;
    const float* beta;
    int32_t pad_out;        /* 1: padded layout with zero halo, 0: compact */
    int32_t fmt16;
    void* out_bf16;
    void* raw_bf16;         /* optional un-normalised 16-bit copy in the output layout (1x1 shortcut operand) */
} smtl_gnapply_args;
int smtl_gnapply_run(const smtl_gnapply_args* a, void* stream);

/* Per-(image, channel) scale / shift of a GroupNorm from the producer-side sums: ss[b, c] = (rstd * gamma, beta -
 * mean * rstd * gamma): what GroupNorm reduces to per image once the producer-side sums are known. */
typedef struct smtl_gnfinalize_args {
    const int64_t* stats;   /* int64 fixed-point cells [stats_replicas, batch, c, 4] (smtl_gemm_args.stats) */
    int32_t stats_replicas, batch, c, groups;
    int64_t pixels;         /* interior pixels per image (h * w) */
    float eps;
    int32_t pad_;
    const float* gamma;
    const float* beta;
    float* ss;              /* fp32 [batch, c, 2] */
} smtl_gnfinalize_args;
int smtl_gnfinalize_run(const smtl_gnfinalize_args* a, void* stream);

/* cudaMemsetAsync on the plan's stream (zeroing the statistics arena before the producers run). */
typedef struct smtl_memset_args {
    void* ptr;
    int64_t bytes;
    int32_t value;
    int32_t pad_;
} smtl_memset_args;
int smtl_memset_run(const smtl_memset_args* a, void* stream);

/* LayerNorm over the channel dim, fp32 or bf16 in, bf16 out; up to two affine outputs from one pass and
 * per-row-group affine parameters (task_norm_{q,k,v}[task], src/util/model.py:133-138).
 * Replaces nn.LayerNorm at src/model/attention.py:338,358,372,494-495,512. */
typedef struct smtl_ln_args {
    const void* x;
    int32_t x_is_bf16;
    int32_t c, ldx;
    int64_t rows;
    float eps;
    int64_t rows_per_group; /* affine set index = row / rows_per_group (use rows for a single set) */
    const float* gamma0;    /* [ngroup, c] */
    const float* beta0;
    void* out0;
    const float* gamma1;    /* optional second affine/output */
    const float* beta1;
    void* out1;
    int32_t ldo;
    int32_t fmt16;          /* format of out0/out1 and of x when x_is_bf16 */
} smtl_ln_args;
int smtl_ln_run(const smtl_ln_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------ data movement
 * Nearest upsample (x2 or explicit size: src = floor(dst * in/out)) fused with the bf16 cast and the zero halo
 * of the following 3x3 conv.  Replaces F.interpolate at src/model/resnet.py:58-61. */
typedef struct smtl_upsample_args {
    const void* x;          /* compact [batch, h, w, c], fp32 or 16-bit (x_fmt16) */
    int32_t batch, h, w, c;
    int32_t oh, ow;
    void* out_bf16;         /* padded layout [batch, oh+2, ow+2, c] */
    int32_t fmt16;
    int32_t x_fmt16;        /* 0: x is fp32; 1: x is 16-bit (fmt16) */
} smtl_upsample_args;
int smtl_upsample_run(const smtl_upsample_args* a, void* stream);

/* Explicit im2col (bf16) for the few convs the shifted-GEMM form does not cover: stride-2 downsamplers
 * (src/model/resnet.py:87,105; diffusers Downsample2D with (0,1,0,1) padding) and tiny-Cin stems (conv_in). */
typedef struct smtl_im2col_args {
    const void* x;          /* compact [batch, h, w, c], fp32 or 16-bit (x_fmt16) */
    int32_t batch, h, w, c;
    int32_t stride, pad_t, pad_l;
    int32_t oh, ow;
    int32_t kpad;           /* row length of the output (>= 9*c, multiple of 64, zero filled) */
    void* out_bf16;         /* [batch*oh*ow, kpad] */
    int32_t fmt16;
    int32_t x_fmt16;        /* 0: x is fp32; 1: x is 16-bit (fmt16; vectorised path only) */
} smtl_im2col_args;
int smtl_im2col_run(const smtl_im2col_args* a, void* stream);

/* [0,255] NCHW rgb -> [-1,1] NHWC fp32 (src/stablemtl_pipeline.py:263). */
typedef struct smtl_rgbprep_args {
    const void* rgb_nchw;   /* float32 or uint8 [batch, 3, h, w] in [0,255] */
    int32_t batch, h, w;
    int32_t src_u8;         /* 0: float32 in [0,255]; 1: uint8; 2: float32 already in [-1,1] (encode_rgb's argument,
                             * src/stablemtl_pipeline.py:607: layout change only) */
    float* out_nhwc;
} smtl_rgbprep_args;
int smtl_rgbprep_run(const smtl_rgbprep_args* a, void* stream);

/* The same normalisation fused with the im2col of the VAE encoder's 3-channel stem conv (diffusers Encoder.conv_in,
 * reached from src/stablemtl_pipeline.py:619): out is the 16-bit GEMM operand [batch*h*w, 64], k = (ky*3 + kx)*3 + channel
 * for k < 27 (pad 1, zero outside the image), zero for k >= 27. */
typedef struct smtl_rgbstem_args {
    const void* rgb_nchw;   /* [batch, 3, h, w] */
    int32_t batch, h, w;
    int32_t src_mode;       /* as smtl_rgbprep_args.src_u8: 0 float32 [0,255], 1 uint8, 2 float32 already in [-1,1] */
    void* out_bf16;
    int32_t fmt16;
    int32_t pad_;
} smtl_rgbstem_args;
int smtl_rgbstem_run(const smtl_rgbstem_args* a, void* stream);

/* UNet input assembly (src/stablemtl_pipeline.py:431-450,557-558,582-584): row group g of the output takes its
 * first 4 channels from latent image first_img[g*images + i], the next 4 from second_img[...], last 4 = 0. */
typedef struct smtl_unetin_args {
    const float* latents;   /* fp32 [n_lat_images, hw, 4] */
    const int32_t* first_img;   /* device int32 [out_images] */
    const int32_t* second_img;  /* device int32 [out_images] */
    int32_t out_images, hw;
    float* out;             /* fp32 [out_images, hw, 12] */
} smtl_unetin_args;
int smtl_unetin_run(const smtl_unetin_args* a, void* stream);

/* Tiny per-pixel channel mix, fp32: y[m, j] = sum_i x[m, i] * w[j, i] + b[j]  (cin, cout <= 16).
 * The 1x1 post_quant_conv of the VAE (diffusers AutoencoderKL, src/stablemtl_pipeline.py:640-642) with the
 * 1/0.18215 latent scaling folded into w. */
typedef struct smtl_chanmix_args {
    const float* x;
    int64_t rows;
    int32_t cin, cout;
    const float* w;         /* [cout, cin] */
    const float* b;         /* [cout] or NULL */
    float* y;
} smtl_chanmix_args;
int smtl_chanmix_run(const smtl_chanmix_args* a, void* stream);

/* Second half of a 3x3 / pad-1 conv with cout <= 4 output channels whose three kx taps were folded into the GEMM's N
 * side (the VAE decoder's conv_out, diffusers Decoder.conv_out reached from src/stablemtl_pipeline.py:643): `partial` is
 * the fp32 output of a 3-segment (ky) implicit GEMM over the PADDED map,
 *     partial[row, kx * cout + co] = sum_ky sum_c a_pad[row + (ky - 1) * (w + 2), c] * W[co, c, ky, kx];
 * this adds the three horizontally shifted entries of every interior pixel and the bias: out[pixel, co], compact fp32. */
typedef struct smtl_headgather_args {
    const float* partial;   /* fp32 [batch * (h+2) * (w+2), ldp] */
    int32_t ldp;
    int32_t batch, h, w;
    int32_t cout;           /* 1..4 */
    const float* bias;      /* fp32 [cout] or NULL */
    float* out;             /* fp32 [batch * h * w, cout] */
} smtl_headgather_args;
int smtl_headgather_run(const smtl_headgather_args* a, void* stream);

/* Task-map epilogue (src/stablemtl_pipeline.py:601,645-654 and the post-processing at :297-366):
 * x is the VAE decoder output fp32 [batch, hw, 3]. */
enum {
    SMTL_MAP_MEAN1 = 0,     /* depth / shading: mean over 3 channels, clip; post = (x+1)/2 */
    SMTL_MAP_RGB3 = 1,      /* albedo: clip; post = (x+1)/2 */
    SMTL_MAP_NORMAL = 2,    /* normal: clip; post = x / max(|x|,0->1) */
    SMTL_MAP_FLOW2 = 3,     /* optical flow: first 2 channels, clip */
    SMTL_MAP_FLOW3 = 4,     /* scene flow: 3 channels, clip */
    SMTL_MAP_SEMANTIC = 5   /* semantic: clip; post = argmin_k |x - palette_k| (first index wins ties) */
};
typedef struct smtl_taskmap_args {
    const float* x;
    int32_t batch, hw, mode;
    float* out_clipped;     /* optional: [batch, cout, hw] planar = single_infer() result */
    float* out_post;        /* optional: [batch, cout, hw] planar post-processed map */
    int64_t* out_ids;       /* semantic only: [batch, hw] */
    const float* palette;   /* semantic only: fp32 [npalette, 3] already mapped to [-1,1] */
    int32_t npalette;
    int32_t pad_;
} smtl_taskmap_args;
int smtl_taskmap_run(const smtl_taskmap_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------ evaluation pre-reductions
 * the evaluation loop that consumes the path's maps reduces every full-resolution map to a few
 * numbers on the host (src/trainer/stablemtl_trainer.py:580-1093).  These two kernels do the data-sized part of
 * that on the device so only the sums cross PCIe.  Both ACCUMULATE into their output: zero it (smtl_memset_run) first. */

/* Sums of the scale/shift least-squares alignment of a predicted map to the ground truth
 * (align_depth_least_square, src/util/alignment.py:122-169: lstsq([pred, 1], gt) over the valid pixels):
 *     sums[b] = { n, sum p, sum g, sum p*p, sum p*g }     fp64, per image
 * scale = (n Spg - Sp Sg) / (n Spp - Sp^2), shift = (Sg - scale Sp) / n on the host. */
typedef struct smtl_lsqsums_args {
    const float* pred;      /* fp32 [batch, hw] */
    const float* gt;        /* fp32 [batch, hw] */
    const uint8_t* valid;   /* [batch, hw] 0 / non-0; NULL = every pixel */
    int32_t batch;
    int32_t pad_;
    int64_t hw;
    double* sums;           /* fp64 [batch, 5] */
} smtl_lsqsums_args;
int smtl_lsqsums_run(const smtl_lsqsums_args* a, void* stream);

/* Confusion-matrix histogram of predicted against true class ids (SemanticMetrics._fast_hist,
 * src/util/metric_semantic.py:34-50): hist[t * n_classes + p] += 1 for every valid pixel with 0 <= t < n_classes
 * (a pixel whose prediction is outside [0, n_classes) is an error: SMTL_EINVAL is NOT raised on the device, the
 * pixel is skipped and counted in hist[n_classes * n_classes]). */
typedef struct smtl_confusion_args {
    const int64_t* label_true;  /* int64 [n] */
    const int64_t* label_pred;  /* int64 [n] */
    const uint8_t* valid;       /* [n] or NULL */
    int64_t n;
    int32_t n_classes;          /* <= 64 */
    int32_t pad_;
    int64_t* hist;              /* int64 [n_classes * n_classes + 1] */
} smtl_confusion_args;
int smtl_confusion_run(const smtl_confusion_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------ plans
 * A plan is an array of (kind, pointer-to-op-struct); smtl_run_plan launches them in order on one stream
 * from native code, so a whole UNet/VAE pass costs one call across the ABI (and can be stream-captured
 * into a CUDA graph by the caller). */
enum {
    SMTL_OP_GEMM = 1, SMTL_OP_FATTN = 2, SMTL_OP_SOFTMAX = 3, SMTL_OP_XATTN = 4, SMTL_OP_TASKATTN = 5,
    /* 6 retired (first-generation two-pass GroupNorm) */ SMTL_OP_LN = 7, SMTL_OP_UPSAMPLE = 8, SMTL_OP_IM2COL = 9, SMTL_OP_RGBPREP = 10,
    SMTL_OP_UNETIN = 11, SMTL_OP_TASKMAP = 12, SMTL_OP_CHANMIX = 13, SMTL_OP_GNAPPLY = 14, SMTL_OP_MEMSET = 15,
    SMTL_OP_GNFINALIZE = 16, SMTL_OP_LSQSUMS = 17, SMTL_OP_CONFUSION = 18, SMTL_OP_RGBSTEM = 19, SMTL_OP_HEADGATHER = 20, SMTL_OP_XATTNF = 21
};
typedef struct smtl_op_ref {
    int32_t kind;
    int32_t pad_;
    const void* op;
} smtl_op_ref;
int smtl_run_plan(const smtl_op_ref* ops, int32_t n_ops, void* stream);
/* number of kernel launches smtl_run_plan(ops) performs (a MEMSET op is a driver memset: 0) */
int smtl_plan_launches(const smtl_op_ref* ops, int32_t n_ops);

/* ------------------------------------------------------------------------------------------------ misc */
int smtl_abi_version(void);
const char* smtl_last_error(void);
/* sizeof() of every struct above, in declaration order, for binding self-checks; returns the count written */
int smtl_struct_sizes(int32_t* out, int32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* STABLEMTL_SM100_H */
