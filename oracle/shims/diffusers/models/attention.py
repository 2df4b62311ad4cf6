"""diffusers.models.attention.{Attention, FeedForward, AdaLayerNorm, BasicTransformerBlock} restated
(0.25.0 semantics).  `Attention` is both the base class of the reference's
SparseCausalAttention (src/model/attention.py:383) and its cross-attention `attn2` (:267)."""
import torch
import torch.nn.functional as F
from torch import nn


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, dropout=0.0, bias=False,
                 upcast_attention=False, upcast_softmax=False, cross_attention_norm=None,
                 cross_attention_norm_num_groups=32, added_kv_proj_dim=None, norm_num_groups=None,
                 spatial_norm_dim=None, out_bias=True, scale_qk=True, only_cross_attention=False, eps=1e-5,
                 rescale_output_factor=1.0, residual_connection=False, _from_deprecated_attn_block=False,
                 processor=None):
        super().__init__()
        self.inner_dim = dim_head * heads
        self.cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.upcast_attention = upcast_attention
        self.upcast_softmax = upcast_softmax
        self.rescale_output_factor = rescale_output_factor
        self.residual_connection = residual_connection
        self.dropout = dropout
        self.scale_qk = scale_qk
        self.scale = dim_head ** -0.5 if scale_qk else 1.0
        self.heads = heads
        self.sliceable_head_dim = heads
        self.added_kv_proj_dim = added_kv_proj_dim
        self.only_cross_attention = only_cross_attention
        self._slice_size = None
        self._use_memory_efficient_attention_xformers = False
        self.group_norm = (nn.GroupNorm(num_channels=query_dim, num_groups=norm_num_groups, eps=eps, affine=True)
                           if norm_num_groups is not None else None)
        self.spatial_norm = None
        self.norm_cross = None
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_v = nn.Linear(self.cross_attention_dim, self.inner_dim, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(dropout)])

    def set_use_memory_efficient_attention_xformers(self, use, attention_op=None):
        self._use_memory_efficient_attention_xformers = use

    def head_to_batch_dim(self, tensor, out_dim=3):
        b, n, d = tensor.shape
        h = self.heads
        tensor = tensor.reshape(b, n, h, d // h).permute(0, 2, 1, 3)
        if out_dim == 3:
            tensor = tensor.reshape(b * h, n, d // h)
        return tensor

    def batch_to_head_dim(self, tensor):
        bh, n, d = tensor.shape
        h = self.heads
        return tensor.reshape(bh // h, h, n, d).permute(0, 2, 1, 3).reshape(bh // h, n, d * h)

    def get_attention_scores(self, query, key, attention_mask=None):
        dtype = query.dtype
        if self.upcast_attention:
            query, key = query.float(), key.float()
        scores = self.scale * torch.bmm(query, key.transpose(-1, -2))
        if attention_mask is not None:
            scores = scores + attention_mask
        if self.upcast_softmax and scores.dtype in (torch.float16, torch.bfloat16):
            scores = scores.float()     # (diffusers upcasts unconditionally; a float64 pinning run must not be downcast)
        return scores.softmax(dim=-1).to(dtype)

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kwargs):
        residual = hidden_states
        input_ndim = hidden_states.ndim
        if input_ndim == 4:
            b, c, h, w = hidden_states.shape
            hidden_states = hidden_states.view(b, c, h * w).transpose(1, 2)
        if self.group_norm is not None:
            hidden_states = self.group_norm(hidden_states.transpose(1, 2)).transpose(1, 2)
        query = self.to_q(hidden_states)
        if encoder_hidden_states is None:
            encoder_hidden_states = hidden_states
        key = self.to_k(encoder_hidden_states)
        value = self.to_v(encoder_hidden_states)
        query, key, value = (self.head_to_batch_dim(t) for t in (query, key, value))
        probs = self.get_attention_scores(query, key, attention_mask)
        hidden_states = self.batch_to_head_dim(torch.bmm(probs, value))
        hidden_states = self.to_out[1](self.to_out[0](hidden_states))
        if input_ndim == 4:
            hidden_states = hidden_states.transpose(-1, -2).reshape(b, c, h, w)
        if self.residual_connection:
            hidden_states = hidden_states + residual
        return hidden_states / self.rescale_output_factor


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, hidden_states):
        hidden_states, gate = self.proj(hidden_states).chunk(2, dim=-1)
        return hidden_states * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", final_dropout=False):
        super().__init__()
        assert activation_fn == "geglu", "only the GEGLU branch is on the StableMTL path"
        inner = int(dim * mult)
        dim_out = dim_out if dim_out is not None else dim
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim_out)])

    def forward(self, hidden_states, scale=1.0):
        for m in self.net:
            hidden_states = m(hidden_states)
        return hidden_states


class AdaLayerNorm(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("AdaLayerNorm is not on the StableMTL path (num_embeds_ada_norm is None)")


class BasicTransformerBlock(nn.Module):
    """Name only: src/util/model.py:5 imports it and never uses it."""
