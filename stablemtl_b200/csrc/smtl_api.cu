// C-ABI glue: error strings, tensor-map encoding through the driver entry point (no link-time libcuda
// dependency, so the library loads on a CPU-only host), struct-size self-check and the native plan runner.
#include "smtl_host.h"

#include <mutex>

namespace smtl_host {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
        else (void)cudaGetLastError();
    });
    return fn;
}

int encode_tmap_bf16_2d(uint64_t out[16], const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver / GPU?)");
        return SMTL_ENODEV;
    }
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || (ld * 2) % 16 != 0 || rows == 0 || cols == 0) {
        set_error("tensor map: base %p / ld %llu not 16-byte aligned or empty extent (%llu x %llu)", base,
                  (unsigned long long)ld, (unsigned long long)rows, (unsigned long long)cols);
        return SMTL_EINVAL;
    }
    alignas(64) CUtensorMap tm;
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu ld %llu box %u x %u)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
        return SMTL_ECUDA;
    }
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    memcpy(out, &tm, 128);
    return SMTL_OK;
}

int num_sms() {
    static std::atomic<int> cache[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    const bool cached = dev >= 0 && dev < 64;
    int n = cached ? cache[dev].load(std::memory_order_relaxed) : 0;
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        if (cached) cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

bool first_use_on_device(std::atomic<uint64_t>& mask) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;     // unknown: always (re)apply
    const uint64_t bit = 1ull << dev;
    return (mask.fetch_or(bit, std::memory_order_acq_rel) & bit) == 0;
}

}  // namespace smtl_host

extern "C" {

int smtl_abi_version(void) { return SMTL_ABI_VERSION; }
const char* smtl_last_error(void) { return smtl_host::get_error(); }

int smtl_struct_sizes(int32_t* out, int32_t cap) {
    const int32_t sizes[] = {
        (int32_t)sizeof(smtl_gemm_seg),     (int32_t)sizeof(smtl_gemm_args),    (int32_t)sizeof(smtl_gemm_op),
        (int32_t)sizeof(smtl_fattn_args),   (int32_t)sizeof(smtl_fattn_op),     (int32_t)sizeof(smtl_softmax_args),
        (int32_t)sizeof(smtl_xattn_args),   (int32_t)sizeof(smtl_xattnf_args),  (int32_t)sizeof(smtl_taskattn_args),
        (int32_t)sizeof(smtl_gnapply_args), (int32_t)sizeof(smtl_gnfinalize_args), (int32_t)sizeof(smtl_memset_args),
        (int32_t)sizeof(smtl_ln_args),      (int32_t)sizeof(smtl_upsample_args), (int32_t)sizeof(smtl_im2col_args),
        (int32_t)sizeof(smtl_rgbprep_args), (int32_t)sizeof(smtl_rgbstem_args), (int32_t)sizeof(smtl_unetin_args),  (int32_t)sizeof(smtl_chanmix_args),
        (int32_t)sizeof(smtl_headgather_args),
        (int32_t)sizeof(smtl_taskmap_args), (int32_t)sizeof(smtl_lsqsums_args), (int32_t)sizeof(smtl_confusion_args),
        (int32_t)sizeof(smtl_op_ref)};
    const int n = (int)(sizeof(sizes) / sizeof(sizes[0]));
    for (int i = 0; i < n && i < cap; ++i) out[i] = sizes[i];
    return n;
}

int smtl_plan_launches(const smtl_op_ref* ops, int32_t n_ops) {
    int n = 0;
    for (int i = 0; i < n_ops; ++i)   // MEMSET is a driver memset, not one of our kernels
        n += (ops[i].kind == SMTL_OP_MEMSET) ? 0 : 1;
    return n;
}

int smtl_run_plan(const smtl_op_ref* ops, int32_t n_ops, void* stream) {
    for (int i = 0; i < n_ops; ++i) {
        int rc;
        const void* p = ops[i].op;
        switch (ops[i].kind) {
            case SMTL_OP_GEMM: rc = smtl_gemm_run((const smtl_gemm_op*)p, stream); break;
            case SMTL_OP_FATTN: rc = smtl_fattn_run((const smtl_fattn_op*)p, stream); break;
            case SMTL_OP_SOFTMAX: rc = smtl_softmax_run((const smtl_softmax_args*)p, stream); break;
            case SMTL_OP_XATTN: rc = smtl_xattn_run((const smtl_xattn_args*)p, stream); break;
            case SMTL_OP_TASKATTN: rc = smtl_taskattn_run((const smtl_taskattn_args*)p, stream); break;
            case SMTL_OP_LN: rc = smtl_ln_run((const smtl_ln_args*)p, stream); break;
            case SMTL_OP_UPSAMPLE: rc = smtl_upsample_run((const smtl_upsample_args*)p, stream); break;
            case SMTL_OP_IM2COL: rc = smtl_im2col_run((const smtl_im2col_args*)p, stream); break;
            case SMTL_OP_RGBPREP: rc = smtl_rgbprep_run((const smtl_rgbprep_args*)p, stream); break;
            case SMTL_OP_UNETIN: rc = smtl_unetin_run((const smtl_unetin_args*)p, stream); break;
            case SMTL_OP_TASKMAP: rc = smtl_taskmap_run((const smtl_taskmap_args*)p, stream); break;
            case SMTL_OP_CHANMIX: rc = smtl_chanmix_run((const smtl_chanmix_args*)p, stream); break;
            case SMTL_OP_GNAPPLY: rc = smtl_gnapply_run((const smtl_gnapply_args*)p, stream); break;
            case SMTL_OP_MEMSET: rc = smtl_memset_run((const smtl_memset_args*)p, stream); break;
            case SMTL_OP_GNFINALIZE: rc = smtl_gnfinalize_run((const smtl_gnfinalize_args*)p, stream); break;
            case SMTL_OP_LSQSUMS: rc = smtl_lsqsums_run((const smtl_lsqsums_args*)p, stream); break;
            case SMTL_OP_CONFUSION: rc = smtl_confusion_run((const smtl_confusion_args*)p, stream); break;
            case SMTL_OP_RGBSTEM: rc = smtl_rgbstem_run((const smtl_rgbstem_args*)p, stream); break;
            case SMTL_OP_HEADGATHER: rc = smtl_headgather_run((const smtl_headgather_args*)p, stream); break;
            case SMTL_OP_XATTNF: rc = smtl_xattnf_run((const smtl_xattnf_args*)p, stream); break;
            default:
                smtl_host::set_error("plan op %d: unknown kind %d", i, ops[i].kind);
                return SMTL_EKIND;
        }
        if (rc != SMTL_OK) return rc;
    }
    return SMTL_OK;
}

}  // extern "C"
