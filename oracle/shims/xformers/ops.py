"""xformers.ops.memory_efficient_attention restated: softmax(q k^T / sqrt(d) + bias) v on [B, M, d] tensors.
Math runs in fp32; the result is returned in the dtype of `q` (the reference feeds fp16, attention.py:392-394)."""
import math

import torch


def memory_efficient_attention(query, key, value, attn_bias=None, p=0.0, scale=None, op=None):
    dt = query.dtype
    q, k, v = query.float(), key.float(), value.float()
    s = scale if scale is not None else 1.0 / math.sqrt(q.shape[-1])
    scores = torch.matmul(q, k.transpose(-1, -2)) * s
    if attn_bias is not None:
        scores = scores + attn_bias.float()
    return torch.matmul(torch.softmax(scores, dim=-1), v).to(dt)
