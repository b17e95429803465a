#!/usr/bin/env python
"""Golden runs of the reference's SAMPLING loop (jacobi_sample_multilevel, lade/decoding.py:137) on CPU with
fixed python / torch seeds.  Output: tests/golden/sample_traces.json.gz.  Re-run: python tests/golden/gen_golden_sample.py"""
from __future__ import annotations

import gzip
import json
import os
import random
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")
from oracle import ref_shim as R  # noqa: E402

TINY = dict(hidden=256, layers=2, heads=2, inter=688, vocab=32000, max_pos=2048)
SMALLV = dict(hidden=256, layers=2, heads=2, inter=688, vocab=96, max_pos=2048)

CASES = [
    # name, model, dtype, W, N, G, pool, P, max_new, wseed, seed, temperature, top_k, top_p, eos
    ("s_fp32_t08_w15n5g15", TINY, "float32", 15, 5, 15, True, 48, 64, 0, 11, 0.8, 0, 1.0, None),
    ("s_fp32_t07_k50_p09_w7n4g7", TINY, "float32", 7, 4, 7, True, 40, 64, 0, 12, 0.7, 50, 0.9, None),
    ("s_fp32_smallv_t1_w5n3g5", SMALLV, "float32", 5, 3, 5, True, 24, 96, 0, 13, 1.0, 0, 1.0, None),
    ("s_fp32_smallv_eos_w6n4g6", SMALLV, "float32", 6, 4, 6, True, 24, 200, 0, 14, 1.0, 0, 1.0, 7),
    ("s_bf16_t08_w15n5g15", TINY, "bfloat16", 15, 5, 15, True, 64, 64, 1, 15, 0.8, 0, 1.0, None),
]


def make_warper(temperature, top_k, top_p):
    from transformers.generation.logits_process import (LogitsProcessorList, TemperatureLogitsWarper,
                                                        TopKLogitsWarper, TopPLogitsWarper)
    lst = LogitsProcessorList()
    if temperature is not None and temperature != 1.0:
        lst.append(TemperatureLogitsWarper(temperature))
    if top_k:
        lst.append(TopKLogitsWarper(top_k=top_k, min_tokens_to_keep=1))
    if top_p is not None and top_p < 1.0:
        lst.append(TopPLogitsWarper(top_p=top_p, min_tokens_to_keep=1))
    return lst


def main():
    from transformers import StoppingCriteriaList, MaxLengthCriteria
    decoding, _ = R.load_reference()
    out = {}
    for (name, mcfg, dt, W, N, G, pool, P, max_new, wseed, seed, temp, top_k, top_p, eos) in CASES:
        cfg = R.make_llama_config(**mcfg)
        model = R.build_reference_model(cfg, seed=wseed, dtype=getattr(torch, dt))
        torch.manual_seed(seed)
        prompt = torch.randint(3, mcfg["vocab"], (1, P))
        decoding.CONFIG_MAP.clear()
        decoding.CONFIG_MAP.update(dict(WINDOW_SIZE=W, LEVEL=N, GUESS_SET_SIZE=G, DEBUG=1, POOL_FROM_PROMPT=int(pool), log=[]))
        random.seed(seed)
        torch.manual_seed(seed + 1000)
        hits_log = []
        with torch.no_grad():
            ids = decoding.jacobi_sample_multilevel(
                model, prompt, logits_warper=make_warper(temp, top_k, top_p),
                stopping_criteria=StoppingCriteriaList([MaxLengthCriteria(P + max_new)]),
                attention_mask=torch.ones_like(prompt), use_cache=True, return_dict_in_generate=False,
                output_attentions=False, output_hidden_states=False, output_scores=False, pad_token_id=0,
                eos_token_id=eos)
        log = decoding.CONFIG_MAP["log"][-1]
        out[name] = dict(model=mcfg, dtype=dt, W=W, N=N, G=G, pool_from_prompt=bool(pool), weight_seed=wseed, seed=seed,
                         prompt=prompt[0].tolist(), max_new=max_new, temperature=temp, top_k=top_k, top_p=top_p,
                         eos_token_id=eos, output_ids=ids[0].tolist(), n_steps=log[1], n_generated=log[0])
        print(f"{name}: {log[0]} tokens in {log[1]} steps")
    with gzip.open(os.path.join(HERE, "sample_traces.json.gz"), "wt") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
