"""The oracle's restatement of the SAMPLING loop (oracle.lookahead.sample_lookahead) reproduces the unmodified
reference's jacobi_sample_multilevel token for token under the same python / torch seeds (CPU)."""
import gzip
import json
import os
import random

import pytest
import torch

from oracle import lookahead as LA
from oracle import llama_ref as LR

GOLD = os.path.join(os.path.dirname(__file__), "golden")
with gzip.open(os.path.join(GOLD, "sample_traces.json.gz"), "rt") as f:
    CASES = json.load(f)


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


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_sampling_matches_reference(name):
    c = CASES[name]
    w = LR.init_weights(c["model"], seed=c["weight_seed"], dtype=getattr(torch, c["dtype"]))
    om = LR.OracleLlama(c["model"], w)
    rng = random.Random(c["seed"])
    torch.manual_seed(c["seed"] + 1000)
    ids, steps = LA.sample_lookahead(c["prompt"], c["max_new"], c["W"], c["N"], c["G"], om,
                                     warper=make_warper(c["temperature"], c["top_k"], c["top_p"]),
                                     pool_from_prompt=c["pool_from_prompt"], eos_token_id=c["eos_token_id"], rng=rng)
    assert ids == c["output_ids"]
    assert steps == c["n_steps"]
