"""Independent fp32 PyTorch LLaMA forward (second opinion for the oracle).

Written from the published LLaMA architecture, not from the oracle: pre-norm RMSNorm(eps 1e-6),
rotary embedding on adjacent pairs (2i, 2i+1) with theta_i = 10000^(-2i/d), causal softmax
attention with 1/sqrt(d) scaling, SwiGLU feed-forward.  Whole-sequence (not incremental).
"""
import numpy as np
import torch


def rmsnorm(x, g, eps=1e-6):
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * g


def rope(x, pos0=0):
    # x: [T, H, D]
    T, H, D = x.shape
    i = torch.arange(0, D, 2, dtype=torch.float64)
    theta = 10000.0 ** (-i / D)
    ang = (torch.arange(T, dtype=torch.float64) + pos0)[:, None] * theta[None, :]
    cos, sin = ang.cos().float()[:, None, :], ang.sin().float()[:, None, :]
    x0, x1 = x[..., 0::2], x[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = x0 * cos - x1 * sin
    out[..., 1::2] = x0 * sin + x1 * cos
    return out


def forward(weights: dict, cfg, tokens):
    """weights: name -> np.ndarray (f16 matrices / f32 gains).  Returns logits [T, V] (float32)."""
    W = {k: torch.from_numpy(np.asarray(v, dtype=np.float32)) for k, v in weights.items()}
    T = len(tokens)
    H, D = cfg.n_head, cfg.head_dim
    x = W["tok_embeddings.weight"][torch.tensor(tokens, dtype=torch.long)]
    mask = torch.full((T, T), float("-inf")).triu(1)
    for l in range(cfg.n_layer):
        p = f"layers.{l}."
        a = rmsnorm(x, W[p + "attention_norm.weight"].reshape(-1))
        q = (a @ W[p + "attention.wq.weight"].T).view(T, H, D)
        k = (a @ W[p + "attention.wk.weight"].T).view(T, H, D)
        v = (a @ W[p + "attention.wv.weight"].T).view(T, H, D)
        q, k = rope(q), rope(k)
        s = torch.einsum("thd,shd->hts", q, k) / (D ** 0.5) + mask
        o = torch.einsum("hts,shd->thd", s.softmax(-1), v).reshape(T, H * D)
        h = x + o @ W[p + "attention.wo.weight"].T
        b = rmsnorm(h, W[p + "ffn_norm.weight"].reshape(-1))
        ff = torch.nn.functional.silu(b @ W[p + "feed_forward.w1.weight"].T) * (b @ W[p + "feed_forward.w3.weight"].T)
        x = h + ff @ W[p + "feed_forward.w2.weight"].T
    x = rmsnorm(x, W["norm.weight"].reshape(-1))
    return (x @ W["output.weight"].T).numpy()
