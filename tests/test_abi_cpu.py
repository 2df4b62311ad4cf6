"""The C-ABI library loads without a GPU, exports every symbol include/stablemtl_sm100.h declares, its structs
match the ctypes mirror, and argument validation fails loudly (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "stablemtl_sm100.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smtl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from stablemtl_b200 import _lib
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)


def test_struct_sizes_and_abi_version():
    from stablemtl_b200 import _lib
    assert _lib.check_struct_sizes()
    src = open(HEADER).read()
    ver = int(re.search(r"#define\s+SMTL_ABI_VERSION\s+(\d+)", src).group(1))
    assert _lib.lib.smtl_abi_version() == ver


def test_header_compiles_as_plain_c(tmp_path):
    """plain pointers and sizes only: the header must be valid C (no C++/torch types in the signatures)."""
    c = tmp_path / "t.c"
    c.write_text('#include "stablemtl_sm100.h"\nint main(void){ smtl_gemm_args a; (void)a; return SMTL_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(c), "-o",
                    str(tmp_path / "t.o")], check=True)


def test_argument_validation_without_gpu():
    from stablemtl_b200 import _lib
    L = _lib
    g, op = L.GemmArgs(), L.GemmOp()
    assert L.lib.smtl_gemm_plan(C.byref(g), C.byref(op)) == -1                 # SMTL_EINVAL: NULL operands
    assert b"NULL" in L.lib.smtl_last_error()
    g.a0, g.b, g.m, g.n, g.k = 256, 256, 128, 64, 64
    assert L.lib.smtl_gemm_plan(C.byref(g), C.byref(op)) == -1                 # no output
    assert b"no output" in L.lib.smtl_last_error()
    g.out_f32 = 256
    g.block_n = 48
    assert L.lib.smtl_gemm_plan(C.byref(g), C.byref(op)) == -1
    assert b"block_n" in L.lib.smtl_last_error()
    g.block_n, g.act, g.n = 0, L.ACT_GEGLU, 320
    assert L.lib.smtl_gemm_plan(C.byref(g), C.byref(op)) == -1                 # GEGLU needs n % 256 == 0
    f, fop = L.FattnArgs(), L.FattnOp()
    assert L.lib.smtl_fattn_plan(C.byref(f), C.byref(fop)) == -1
    with pytest.raises(L.SmtlError):
        L.check(-1, "x")
    ref = (L.OpRef * 1)()
    ref[0].kind = 99
    assert L.lib.smtl_plan_launches(ref, 1) < 0 or L.lib.smtl_run_plan(ref, 1, None) == -4   # SMTL_EKIND


def test_no_cpu_fallback():
    """The product path refuses to run without a CUDA device instead of silently using torch or the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from stablemtl_b200 import synth
    from stablemtl_b200.pipeline import StableMTLEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        StableMTLEngine(synth.TINY_UNET, synth.TINY_VAE, {}, {}, {})


def test_product_path_never_imports_the_oracle():
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "stablemtl_b200")):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, fn)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M) or "/root/reference" in s.replace(
                        "under /root/reference", ""):
                    bad.append(fn)
    assert not bad, bad
