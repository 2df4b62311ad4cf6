"""SASS audit of the built library (cuobjdump, no GPU needed): the tensor-core kernels really are tcgen05 + TMA code for
sm_100a, and the issue path has no uniform-register waterfall loops.

The second point guards the biggest single finding of round 1: when `tcgen05.mma` / `tcgen05.commit` /
`cp.async.bulk.tensor` sit under a lane-id branch, ptxas wraps each of them in an ELECT / R2UR.BROADCAST / BRA.U.ANY
loop (~165 clk per MMA).  With a converged warp and `elect.sync` at the instruction there is no BRA.U.ANY at all."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "stablemtl_b200", "libstablemtl_sm100.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None or not os.path.exists(LIB):
        pytest.skip("cuobjdump or the built library is not available")
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, timeout=600).stdout
    funcs, name = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            funcs[name] = []
        elif name is not None:
            funcs[name].append(line)
    assert "arch = sm_100a" in out
    return funcs


def _count(lines, mnemonic):
    return sum(1 for l in lines if mnemonic in l)


def test_tensor_kernels_are_tcgen05_and_tma(sass):
    gemm = {k: v for k, v in sass.items() if "smtl_gemm_kernel" in k or "smtl_gemmT_kernel" in k}
    attn = {k: v for k, v in sass.items() if "smtl_fattn4_kernel" in k or "smtl_vattn_kernel" in k}
    assert len(gemm) >= 15 and len(attn) >= 3                 # two formats of the d = 64 kernel + the d = 512 one
    for name, lines in list(gemm.items()) + list(attn.items()):
        assert _count(lines, "UTCHMMA") >= 4, name            # tcgen05.mma
        assert _count(lines, "UTMALDG") >= 2, name            # cp.async.bulk.tensor
        assert _count(lines, "UTCBAR") >= 2, name             # tcgen05.commit -> mbarrier
        assert _count(lines, "LDTM") >= 1, name               # tcgen05.ld in the epilogue
        assert _count(lines, " HMMA.") == 0, name             # no mma.sync fallback (UTCHMMA.2CTA also contains "HMMA.")
    pair = [k for k in gemm if "ELi2E" in k]
    assert pair and all(_count(gemm[k], "UTCHMMA.2CTA") >= 4 for k in pair)      # cta_group::2


def test_no_uniform_register_waterfall_on_the_issue_path(sass):
    for name, lines in sass.items():
        if "smtl_gemm_kernel" in name or "smtl_gemmT_kernel" in name or "smtl_fattn4_kernel" in name or "smtl_vattn_kernel" in name:
            assert _count(lines, "BRA.U.ANY") == 0, f"{name}: waterfall loop around a uniform-datapath instruction"
