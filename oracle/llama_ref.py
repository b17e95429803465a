"""TEST INFRASTRUCTURE ONLY -- torch restatement of the floating-point side of lade's step.

Functional Llama forward over the lookahead step rows, following the reference's eager path
(``/root/reference/lade/models/modeling_llama.py``): RMSNorm ``:222-227``, rotary tables
``:240-266`` and application ``:342-346``, attention ``:492-558`` (QK^T, ``/ sqrt(d)`` as a
division ``:523``, additive ``finfo.min`` mask ``:536``, fp32 softmax cast back ``:539``, PV
``:541``), SwiGLU MLP ``:378``, decoder layer ``:858-889``, final norm + lm_head ``:1240,1541-1544``.

It is the checker for the CUDA kernels (same rounding points as the reference, so it is the "torch
reference" of the floating-point kernels) and, timed on the host cores, the ``cpu_baseline`` /
``--impl reference`` leg of ``bench.py``.  The product package never imports it.

Parity status: PINNED through ``tests/golden`` (per-step argmax tokens and final ids of the
unmodified reference model run from /root/reference with the same seeded weights).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch

from . import lookahead as LA


def init_weights(cfg: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random-init Llama weights (normal(0, 0.02), norms = 1), name-sorted draw order.

    Identical draws to ``oracle.ref_shim.build_reference_model`` so the unmodified reference model
    and this functional restatement share weights bit-for-bit.
    """
    H, L, I, V = cfg["hidden"], cfg["layers"], cfg["inter"], cfg["vocab"]
    nh, nkv = cfg["heads"], cfg.get("kv_heads") or cfg["heads"]
    D = H // nh
    shapes = {"lm_head.weight": (V, H), "model.embed_tokens.weight": (V, H), "model.norm.weight": (H,)}
    for i in range(L):
        p = f"model.layers.{i}."
        shapes[p + "input_layernorm.weight"] = (H,)
        shapes[p + "post_attention_layernorm.weight"] = (H,)
        shapes[p + "self_attn.q_proj.weight"] = (nh * D, H)
        shapes[p + "self_attn.k_proj.weight"] = (nkv * D, H)
        shapes[p + "self_attn.v_proj.weight"] = (nkv * D, H)
        shapes[p + "self_attn.o_proj.weight"] = (H, nh * D)
        shapes[p + "mlp.gate_proj.weight"] = (I, H)
        shapes[p + "mlp.up_proj.weight"] = (I, H)
        shapes[p + "mlp.down_proj.weight"] = (H, I)
    g = torch.Generator().manual_seed(seed)
    w = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if len(shp) >= 2:
            w[name] = (torch.randn(shp, generator=g) * 0.02).to(dtype)
        else:
            w[name] = torch.ones(shp, dtype=dtype)
    return w


def rope_tables(D: int, max_pos: int, theta: float, dtype, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin caches as modeling_llama.py:240-256 (fp32 math, then cast to the model dtype :264-265)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, D, 2).float() / D))
    t = torch.arange(max_pos, dtype=inv_freq.dtype)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(dtype).to(device), emb.sin().to(dtype).to(device)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return w * xf.to(dt)


def rotate_half(x):
    x1 = x[..., : x.shape[-1] // 2]
    x2 = x[..., x.shape[-1] // 2:]
    return torch.cat((-x2, x1), dim=-1)


def additive_mask(vis: torch.Tensor, kv_len: int, dtype) -> torch.Tensor:
    """[q, kv_len+q] additive mask: 0 where visible, finfo(dtype).min elsewhere (modeling :122,205-206)."""
    q = vis.shape[0]
    m = torch.full((q, q), torch.finfo(dtype).min, dtype=dtype, device=vis.device)
    m.masked_fill_(vis, 0)
    if kv_len > 0:
        m = torch.cat([torch.zeros(q, kv_len, dtype=dtype, device=vis.device), m], dim=-1)
    return m


def eager_attention(q, k, v, mask, n_rep: int = 1):
    """Reference eager attention numerics on [H, q, D] x [Hkv, kv, D] (modeling_llama.py:520-541)."""
    if n_rep > 1:
        k = k.repeat_interleave(n_rep, dim=0)
        v = v.repeat_interleave(n_rep, dim=0)
    D = q.shape[-1]
    s = torch.matmul(q, k.transpose(1, 2)) / math.sqrt(D)
    s = s + mask
    p = torch.nn.functional.softmax(s, dim=-1, dtype=torch.float32).to(q.dtype)
    return torch.matmul(p, v)


class OracleLlama:
    """Functional Llama with a growing KV cache, driven by ``oracle.lookahead.greedy_lookahead``."""

    def __init__(self, cfg: dict, weights: Dict[str, torch.Tensor], device="cpu"):
        self.cfg = cfg
        self.dev = torch.device(device)
        self.w = {k: v.to(self.dev) for k, v in weights.items()}
        self.dtype = self.w["lm_head.weight"].dtype
        self.H, self.L = cfg["hidden"], cfg["layers"]
        self.nh = cfg["heads"]
        self.nkv = cfg.get("kv_heads") or cfg["heads"]
        self.D = self.H // self.nh
        self.eps = cfg.get("eps", 1e-5)
        self.cos, self.sin = rope_tables(self.D, cfg.get("max_pos", 4096), cfg.get("rope_theta", 10000.0),
                                         self.dtype, self.dev)
        self.reset()

    def reset(self):
        self.k_cache: List[Optional[torch.Tensor]] = [None] * self.L
        self.v_cache: List[Optional[torch.Tensor]] = [None] * self.L
        self.last_hidden = None
        self.last_attn_io = None

    # -- one forward over the step rows, returns fp32 logits [q, V] -------------------------------
    def forward_rows(self, ids, pos, vis_mask: torch.Tensor, kv_len: int, capture_layer: int = -1):
        w, dt = self.w, self.dtype
        ids_t = torch.as_tensor(ids, dtype=torch.long, device=self.dev)
        pos_t = torch.as_tensor(pos, dtype=torch.long, device=self.dev)
        q_len = ids_t.numel()
        h = w["model.embed_tokens.weight"][ids_t]
        mask = additive_mask(vis_mask.to(self.dev), kv_len, dt)
        cos, sin = self.cos[pos_t], self.sin[pos_t]
        for i in range(self.L):
            p = f"model.layers.{i}."
            res = h
            x = rms_norm(h, w[p + "input_layernorm.weight"], self.eps)
            q = torch.nn.functional.linear(x, w[p + "self_attn.q_proj.weight"]).view(q_len, self.nh, self.D).transpose(0, 1)
            k = torch.nn.functional.linear(x, w[p + "self_attn.k_proj.weight"]).view(q_len, self.nkv, self.D).transpose(0, 1)
            v = torch.nn.functional.linear(x, w[p + "self_attn.v_proj.weight"]).view(q_len, self.nkv, self.D).transpose(0, 1)
            q = (q * cos) + (rotate_half(q) * sin)
            k = (k * cos) + (rotate_half(k) * sin)
            if self.k_cache[i] is not None and kv_len > 0:
                k = torch.cat([self.k_cache[i][:, :kv_len], k], dim=1)
                v = torch.cat([self.v_cache[i][:, :kv_len], v], dim=1)
            self.k_cache[i], self.v_cache[i] = k, v
            o = eager_attention(q, k, v, mask, self.nh // self.nkv)
            if i == capture_layer:
                self.last_attn_io = (q.clone(), k.clone(), v.clone(), o.clone())
            o = o.transpose(0, 1).reshape(q_len, self.nh * self.D)
            h = res + torch.nn.functional.linear(o, w[p + "self_attn.o_proj.weight"])
            res = h
            x = rms_norm(h, w[p + "post_attention_layernorm.weight"], self.eps)
            g = torch.nn.functional.linear(x, w[p + "mlp.gate_proj.weight"])
            u = torch.nn.functional.linear(x, w[p + "mlp.up_proj.weight"])
            h = res + torch.nn.functional.linear(torch.nn.functional.silu(g) * u, w[p + "mlp.down_proj.weight"])
        h = rms_norm(h, w["model.norm.weight"], self.eps)
        self.last_hidden = h
        return torch.nn.functional.linear(h, w["lm_head.weight"]).float()

    # -- StepFn / CompactFn for oracle.lookahead.greedy_lookahead ---------------------------------
    def step_fn(self, lay: LA.StepLayout, kv_len: int):
        vis = torch.from_numpy(LA.step_mask(lay))
        logits = self.forward_rows(lay.ids, lay.pos, vis, kv_len)
        self.last_logits = logits
        q = lay.q_len
        lg = lay.n_guess_tok
        window = lay.level_sizes[-1]
        out_tok = int(torch.argmax(logits[lay.n_input - 1]))
        inp = torch.argmax(logits[q - lg - window:q - lg], dim=-1).tolist()
        guess = torch.argmax(logits[q - lg:], dim=-1).tolist() if lg > 0 else []
        return out_tok, inp, guess

    def compact_fn(self, dst: int, src: int, n: int, new_len: int):
        for i in range(self.L):
            if n > 0:
                self.k_cache[i][:, dst:dst + n] = self.k_cache[i][:, src:src + n].clone()
                self.v_cache[i][:, dst:dst + n] = self.v_cache[i][:, src:src + n].clone()
            self.k_cache[i] = self.k_cache[i][:, :new_len]
            self.v_cache[i] = self.v_cache[i][:, :new_len]

    # -- plain autoregressive greedy (comparator) --------------------------------
    def plain_greedy(self, prompt, max_new: int):
        self.reset()
        ids = list(prompt)
        kv = 0
        feed = ids
        for _ in range(max_new):
            n = len(feed)
            vis = torch.tril(torch.ones(n, n, dtype=torch.bool))
            logits = self.forward_rows(feed, list(range(kv, kv + n)), vis, kv)
            kv += n
            nxt = int(torch.argmax(logits[-1]))
            ids.append(nxt)
            feed = [nxt]
        return ids
