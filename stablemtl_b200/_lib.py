"""ctypes binding of libstablemtl_sm100.so (see include/stablemtl_sm100.h).

The structures below mirror the C header field by field; `check_struct_sizes()` compares `ctypes.sizeof`
with the sizes the library reports so a drifted binding fails loudly instead of corrupting arguments.
There is deliberately NO fallback: if the shared library is missing the import raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstablemtl_sm100.so")

MAX_SEG = 12
MAX_TASKS = 8
MAX_XATTN_TOKENS = 8

ACT_NONE, ACT_GELU, ACT_GEGLU, ACT_SILU = 0, 1, 2, 3
(ROWMAP_IDENTITY, ROWMAP_CONV_PAD, ROWMAP_CONV_PAD_UP2, ROWMAP_PAD_KEEP, ROWMAP_TO_PAD,
 ROWMAP_UP2_PAD) = range(6)
FMT_BF16, FMT_F16 = 0, 1
MAP_MEAN1, MAP_RGB3, MAP_NORMAL, MAP_FLOW2, MAP_FLOW3, MAP_SEMANTIC = range(6)
(OP_GEMM, OP_FATTN, OP_SOFTMAX, OP_XATTN, OP_TASKATTN, _OP_RETIRED6, OP_LN, OP_UPSAMPLE, OP_IM2COL, OP_RGBPREP,
 OP_UNETIN, OP_TASKMAP, OP_CHANMIX, OP_GNAPPLY, OP_MEMSET, OP_GNFINALIZE, OP_LSQSUMS, OP_CONFUSION, OP_RGBSTEM, OP_HEADGATHER, OP_XATTNF) = range(1, 22)

vp = C.c_void_p
i32 = C.c_int32
i64 = C.c_int64
f32 = C.c_float


class GemmSeg(C.Structure):
    _fields_ = [("row_shift", i32), ("kblocks", i32), ("src", i32), ("a_col0", i32)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("a0", vp), ("a1", vp), ("b", vp),
        ("a0_rows", i64), ("a1_rows", i64),
        ("a0_cols", i32), ("a1_cols", i32),
        ("a0_ld", i32), ("a1_ld", i32),
        ("m", i64),
        ("n", i32), ("k", i32), ("ldb", i32),
        ("nseg", i32),
        ("seg", GemmSeg * MAX_SEG),
        ("bias", vp),
        ("bias_per_row", i32), ("act", i32),
        ("res1", vp), ("res2", vp),
        ("ldres", i32),
        ("out_f32", vp), ("out_bf16", vp), ("aux_bf16", vp),
        ("ldc", i32), ("ld_aux", i32),
        ("rowmap", i32),
        ("img_h", i32), ("img_w", i32),
        ("block_n", i32), ("fmt16", i32),
        ("res_fmt16", i32), ("stats_replicas", i32),
        ("stats", vp),
        ("stats_rows_per_image", i32), ("stats_images", i32),
        ("cta_group", i32), ("up_parity", i32),
        ("group_rows", i64),
        ("tile_order", i32), ("stats_group", i32),
    ]


class GemmOp(C.Structure):
    _fields_ = [
        ("args", GemmArgs),
        ("tmap_a0", C.c_uint64 * 16), ("tmap_a1", C.c_uint64 * 16), ("tmap_b", C.c_uint64 * 16),
        ("block_n", i32), ("grid", i32), ("tiles_m", i32), ("tiles_n", i32),
        ("total_kblocks", i32), ("smem_bytes", i32), ("cta_group", i32), ("grouped", i32),
        ("ngrp", i32), ("sp", i32), ("sw", i32),
        ("grp", (i32 * 6) * MAX_SEG),
        ("tmap_x8", (C.c_uint64 * 16) * 2),
        ("tile_rpi", i64), ("tiles_per_img", i32), ("pair_split", i32),
    ]


class FattnArgs(C.Structure):
    _fields_ = [
        ("qkv", vp), ("ld", i32), ("q_col0", i32), ("k_col0", i32), ("v_col0", i32),
        ("batch", i32), ("ntok", i32), ("heads", i32),
        ("out_bf16", vp), ("ldo", i32), ("scale", f32), ("fmt16", i32), ("head_dim", i32),
    ]


class FattnOp(C.Structure):
    _fields_ = [("args", FattnArgs), ("tmap_qkv", C.c_uint64 * 16), ("tmap_kv", C.c_uint64 * 16), ("grid_x", i32), ("grid_y", i32),
                ("smem_bytes", i32), ("pad_", i32)]


class SoftmaxArgs(C.Structure):
    _fields_ = [("s", vp), ("rows", i64), ("n", i32), ("lds", i32), ("scale", f32), ("p_bf16", vp), ("ldp", i32), ("fmt16", i32)]


class XattnArgs(C.Structure):
    _fields_ = [
        ("q_bf16", vp), ("ldq", i32), ("rows", i64), ("heads", i32),
        ("kc", vp), ("vc", vp),
        ("ntok", i32 * MAX_TASKS), ("task_of_group", i32 * MAX_TASKS),
        ("rows_per_group", i64),
        ("out_bf16", vp), ("ldo", i32), ("scale", f32), ("fmt16", i32), ("ntok_pad", i32),
    ]


class XattnFArgs(C.Structure):
    _fields_ = [
        ("hs", vp), ("ldh", i32), ("heads", i32), ("rows", i64), ("rows_per_group", i64),
        ("task_of_group", i32 * MAX_TASKS), ("ntok_pad", i32), ("fmt16", i32),
        ("ap", vp), ("ca", vp), ("bmt", vp), ("bo", vp), ("gamma3", vp), ("beta3", vp),
        ("out_bf16", vp), ("ldo", i32), ("eps2", f32), ("eps3", f32), ("pad_", i32),
    ]


class TaskAttnArgs(C.Structure):
    _fields_ = [
        ("q_bf16", vp), ("k_bf16", vp), ("v_bf16", vp), ("out_bf16", vp),
        ("c", i32), ("nheads", i32), ("n_main", i32), ("n_src", i32),
        ("rows_per_group", i64),
        ("main_task", i32 * MAX_TASKS), ("src_task", i32 * MAX_TASKS),
        ("exclude_self", i32), ("scale", f32), ("fmt16", i32), ("pad_", i32),
    ]


class GnApplyArgs(C.Structure):
    _fields_ = [
        ("x0", vp), ("x1", vp), ("c0", i32), ("c1", i32),
        ("x_fmt16", i32), ("stats_replicas", i32), ("x_padded", i32), ("pad2_", i32),
        ("stats0", vp), ("stats1", vp),
        ("batch", i32), ("h", i32), ("w", i32), ("groups", i32), ("eps", f32), ("silu", i32),
        ("gamma", vp), ("beta", vp),
        ("pad_out", i32), ("fmt16", i32),
        ("out_bf16", vp), ("raw_bf16", vp),
    ]


class GnFinalizeArgs(C.Structure):
    _fields_ = [("stats", vp), ("stats_replicas", i32), ("batch", i32), ("c", i32), ("groups", i32), ("pixels", i64),
                ("eps", f32), ("pad_", i32), ("gamma", vp), ("beta", vp), ("ss", vp)]


class MemsetArgs(C.Structure):
    _fields_ = [("ptr", vp), ("bytes", i64), ("value", i32), ("pad_", i32)]


class LnArgs(C.Structure):
    _fields_ = [
        ("x", vp), ("x_is_bf16", i32), ("c", i32), ("ldx", i32),
        ("rows", i64), ("eps", f32), ("rows_per_group", i64),
        ("gamma0", vp), ("beta0", vp), ("out0", vp),
        ("gamma1", vp), ("beta1", vp), ("out1", vp),
        ("ldo", i32), ("fmt16", i32),
    ]


class UpsampleArgs(C.Structure):
    _fields_ = [("x", vp), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("oh", i32), ("ow", i32),
                ("out_bf16", vp), ("fmt16", i32), ("x_fmt16", i32)]


class Im2colArgs(C.Structure):
    _fields_ = [("x", vp), ("batch", i32), ("h", i32), ("w", i32), ("c", i32), ("stride", i32), ("pad_t", i32),
                ("pad_l", i32), ("oh", i32), ("ow", i32), ("kpad", i32), ("out_bf16", vp), ("fmt16", i32), ("x_fmt16", i32)]


class RgbprepArgs(C.Structure):
    _fields_ = [("rgb_nchw", vp), ("batch", i32), ("h", i32), ("w", i32), ("src_u8", i32), ("out_nhwc", vp)]


class RgbstemArgs(C.Structure):
    _fields_ = [("rgb_nchw", vp), ("batch", i32), ("h", i32), ("w", i32), ("src_mode", i32), ("out_bf16", vp),
                ("fmt16", i32), ("pad_", i32)]


class UnetinArgs(C.Structure):
    _fields_ = [("latents", vp), ("first_img", vp), ("second_img", vp), ("out_images", i32), ("hw", i32),
                ("out", vp)]


class ChanmixArgs(C.Structure):
    _fields_ = [("x", vp), ("rows", i64), ("cin", i32), ("cout", i32), ("w", vp), ("b", vp), ("y", vp)]


class HeadGatherArgs(C.Structure):
    _fields_ = [("partial", vp), ("ldp", i32), ("batch", i32), ("h", i32), ("w", i32), ("cout", i32), ("bias", vp),
                ("out", vp)]


class TaskmapArgs(C.Structure):
    _fields_ = [("x", vp), ("batch", i32), ("hw", i32), ("mode", i32), ("out_clipped", vp), ("out_post", vp),
                ("out_ids", vp), ("palette", vp), ("npalette", i32), ("pad_", i32)]


class LsqSumsArgs(C.Structure):
    _fields_ = [("pred", vp), ("gt", vp), ("valid", vp), ("batch", i32), ("pad_", i32), ("hw", i64), ("sums", vp)]


class ConfusionArgs(C.Structure):
    _fields_ = [("label_true", vp), ("label_pred", vp), ("valid", vp), ("n", i64), ("n_classes", i32), ("pad_", i32),
                ("hist", vp)]


class OpRef(C.Structure):
    _fields_ = [("kind", i32), ("pad_", i32), ("op", vp)]


STRUCTS_IN_HEADER_ORDER = [GemmSeg, GemmArgs, GemmOp, FattnArgs, FattnOp, SoftmaxArgs, XattnArgs, XattnFArgs, TaskAttnArgs,
                           GnApplyArgs, GnFinalizeArgs, MemsetArgs, LnArgs, UpsampleArgs, Im2colArgs, RgbprepArgs, RgbstemArgs, UnetinArgs, ChanmixArgs, HeadGatherArgs, TaskmapArgs, LsqSumsArgs, ConfusionArgs, OpRef]

EXPORTS = [
    "smtl_gemm_plan", "smtl_gemm_run", "smtl_fattn_plan", "smtl_fattn_run", "smtl_softmax_run", "smtl_xattn_run", "smtl_xattnf_run", "smtl_xattnf_supported",
    "smtl_taskattn_run", "smtl_gnapply_run", "smtl_gnfinalize_run", "smtl_memset_run", "smtl_ln_run", "smtl_upsample_run", "smtl_im2col_run", "smtl_rgbprep_run", "smtl_rgbstem_run",
    "smtl_unetin_run", "smtl_chanmix_run", "smtl_headgather_run", "smtl_taskmap_run", "smtl_lsqsums_run", "smtl_confusion_run", "smtl_run_plan", "smtl_plan_launches", "smtl_abi_version",
    "smtl_last_error", "smtl_struct_sizes",
]


class SmtlError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU or PyTorch fallback for the CUDA path)")
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        getattr(lib, name)  # raises AttributeError if the symbol is not exported
    lib.smtl_last_error.restype = C.c_char_p
    lib.smtl_abi_version.restype = C.c_int
    for name in EXPORTS:
        if name.endswith("_run") or name.endswith("_plan"):
            getattr(lib, name).restype = C.c_int
    lib.smtl_run_plan.argtypes = [C.POINTER(OpRef), i32, vp]
    lib.smtl_plan_launches.argtypes = [C.POINTER(OpRef), i32]
    lib.smtl_plan_launches.restype = C.c_int
    lib.smtl_struct_sizes.argtypes = [C.POINTER(i32), i32]
    lib.smtl_struct_sizes.restype = C.c_int
    return lib


lib = _load()


def check(rc, what=""):
    if rc != 0:
        raise SmtlError(f"{what} failed with code {rc}: {lib.smtl_last_error().decode()}")


def check_struct_sizes():
    n = len(STRUCTS_IN_HEADER_ORDER)
    buf = (i32 * 64)()
    got = lib.smtl_struct_sizes(buf, 64)
    if got != n:
        raise SmtlError(f"binding knows {n} structs, library reports {got}")
    for i, st in enumerate(STRUCTS_IN_HEADER_ORDER):
        if C.sizeof(st) != buf[i]:
            raise SmtlError(f"sizeof({st.__name__}) = {C.sizeof(st)} in the binding but {buf[i]} in the library")
    return True


check_struct_sizes()
