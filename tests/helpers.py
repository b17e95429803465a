"""Shared helpers for the parity tests."""
import ctypes as C
import gzip
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_cases():
    with gzip.open(os.path.join(GOLD, "greedy_traces.json.gz"), "rt") as f:
        return json.load(f)


def rows_to_bool(rows):
    return np.array([[ch == "1" for ch in r] for r in rows], dtype=bool)


def build_hf_llama(model_cfg: dict, weight_seed: int, dtype=torch.bfloat16, device="cuda"):
    """HF LlamaForCausalLM with the oracle's seeded weights (same draws as the reference fixtures)."""
    from transformers import LlamaConfig, LlamaForCausalLM
    from oracle import llama_ref as LR

    cfg = LlamaConfig(
        hidden_size=model_cfg["hidden"], num_hidden_layers=model_cfg["layers"],
        num_attention_heads=model_cfg["heads"], num_key_value_heads=model_cfg.get("kv_heads") or model_cfg["heads"],
        intermediate_size=model_cfg["inter"], vocab_size=model_cfg["vocab"],
        max_position_embeddings=model_cfg.get("max_pos", 2048), rms_norm_eps=model_cfg.get("eps", 1e-5),
        tie_word_embeddings=False, attention_bias=False, hidden_act="silu",
        rope_parameters={"rope_type": "default", "rope_theta": model_cfg.get("rope_theta", 10000.0)},
    )
    with torch.device("meta"):
        model = LlamaForCausalLM(cfg)
    w = LR.init_weights(model_cfg, seed=weight_seed, dtype=dtype)
    model = model.to_empty(device=device)
    missing = model.load_state_dict({k: v.to(device) for k, v in w.items()}, strict=False)
    assert not [k for k in missing.missing_keys if "rotary" not in k and "inv_freq" not in k], missing
    model = model.to(dtype)
    # buffers (rotary inv_freq) are not part of the state dict: rebuild
    if hasattr(model.model, "rotary_emb"):
        re = model.model.rotary_emb
        D = cfg.hidden_size // cfg.num_attention_heads
        inv = 1.0 / (model_cfg.get("rope_theta", 10000.0) ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
        re.inv_freq = inv.to(device)
        if hasattr(re, "original_inv_freq"):
            re.original_inv_freq = inv.to(device)
    model.eval()
    return model, w


def make_lade_config(W, N, G, vocab, cap, pool=False, eos=()):
    from lookaheaddecoding_b200._cabi import LadeConfig
    c = LadeConfig()
    c.window_size, c.level, c.guess_set_size = W, N, G
    c.pool_from_prompt = int(pool)
    c.vocab_size = vocab
    c.max_total_len = cap
    c.n_eos = len(eos)
    for i, e in enumerate(eos):
        c.eos_token_id[i] = e
    c.dist_workers, c.rank = 1, 0
    return c
