"""xformers.ops.memory_efficient_attention restated: softmax(q k^T / sqrt(d) + bias) v on [B, M, d] tensors.
Math runs in fp32; the result is returned in the dtype of `q` (the reference feeds fp16, attention.py:392-394)."""
import math
import os

import torch


def memory_efficient_attention(query, key, value, attn_bias=None, p=0.0, scale=None, op=None):
    dt = query.dtype
    ct = torch.float64 if os.environ.get("ORACLE_SHIM_FP64") == "1" else torch.float32   # fp64 only while pinning
    q, k, v = query.to(ct), key.to(ct), value.to(ct)
    s = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    scores = torch.matmul(q, k.transpose(-1, -2)) * s
    if attn_bias is not None:
        scores = scores + attn_bias.to(ct)
    return torch.matmul(torch.softmax(scores, dim=-1), v).to(dt)
