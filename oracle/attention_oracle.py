"""CPU oracle for `StandardAttention` (TEST INFRASTRUCTURE - never on the product path).

Plain-PyTorch restatement of models/ADNssd.py:38-47 (to_qkv -> per-head softmax(q k^T * dim_head^-0.5) v -> to_out),
differentiable by autograd (run it in float64).  Parameter names are the module's state_dict keys.  Pinned against the
unmodified reference class in tests/test_oracle_vs_golden.py::test_attention_oracle_matches_reference."""
import torch


def sdpa_packed(qkv, heads, dim_head):
    """qkv (B, L, 3 * heads * dim_head) in the reference's '(h d)' order -> (B, L, heads * dim_head)   (:41-46)."""
    B, L, _ = qkv.shape
    q, k, v = (t.reshape(B, L, heads, dim_head).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1))
    dots = torch.einsum("bhid,bhjd->bhij", q, k) * dim_head ** -0.5
    out = torch.einsum("bhij,bhjd->bhid", dots.softmax(dim=-1), v)
    return out.permute(0, 2, 1, 3).reshape(B, L, heads * dim_head)


def attention_forward(p, x, heads, dim_head):
    """models/ADNssd.py:38-47; p: 'to_qkv.weight' (3 * inner, dim), 'to_out.weight' (dim, inner), 'to_out.bias' (dim)."""
    qkv = x @ p["to_qkv.weight"].t()
    return sdpa_packed(qkv, heads, dim_head) @ p["to_out.weight"].t() + p["to_out.bias"]


def init_params(dim, heads, dim_head, seed=0, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    inner = heads * dim_head
    return {"to_qkv.weight": (torch.randn(3 * inner, dim, generator=g, dtype=torch.float64) * (1.5 / dim ** 0.5)).to(dtype),
            "to_out.weight": (torch.randn(dim, inner, generator=g, dtype=torch.float64) / inner ** 0.5).to(dtype),
            "to_out.bias": (torch.randn(dim, generator=g, dtype=torch.float64) * 0.2).to(dtype)}
