"""Weight ingest: reference state-dict tensors -> the layouts the sm_100a kernels consume.

Everything here is load-time host work (reference src/util/model.py:196-230, src/model/unet.py:447-481):
bf16 casts, KRSC conv filters, GEGLU tile interleave and constant folding (time embedding, text keys/values).
"""
import torch


def interleave_geglu(w, bias=None, tile=256):
    """diffusers GEGLU proj: rows [0,4C) are the value half, [4C,8C) the gate half (Appendix A of SURVEY.md).
    The GEMM epilogue wants both halves of an output column inside one `tile`-row B tile: [value 128 | gate 128]."""
    n2 = w.shape[0] // 2
    half = tile // 2
    assert n2 % half == 0, f"GEGLU inner dim {n2} not a multiple of {half}"
    v = w[:n2].reshape(n2 // half, half, -1)
    g = w[n2:].reshape(n2 // half, half, -1)
    wi = torch.cat([v, g], dim=1).reshape(2 * n2, -1).contiguous()
    bi = None
    if bias is not None:
        bv = bias[:n2].reshape(n2 // half, half)
        bg = bias[n2:].reshape(n2 // half, half)
        bi = torch.cat([bv, bg], dim=1).reshape(2 * n2).contiguous()
    return wi, bi


def conv_weight_matrix(w):
    """[Cout, Cin, 3, 3] -> [Cout, tap*Cin + c] with tap = ky*3 + kx (KRSC)."""
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).contiguous()
