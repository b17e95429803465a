"""Sampling lookahead path (jacobi_sample_multilevel) on the CUDA engine.

* top_k=1 degenerates to greedy: the sampling machinery (rejection test with prob in {0,1}, residual draw,
  commit of an external decision, KV compaction) must reproduce the greedy ids;
* fixed seeds: the engine's sampled ids equal the oracle restatement of the reference loop run on the same
  device with the same python/torch RNG seeds (the oracle itself is pinned to the unmodified reference's
  sampled ids in tests/test_oracle_sampling.py).  Logits differ in the last bf16 bits between the two, so a
  draw can flip only where a CDF boundary falls inside that noise; the test allows one such divergence per
  case provided the two candidate tokens both carry probability mass (> 1e-4) at that position;
* plugin surface: model.generate(do_sample=True, temperature=...) routes to the sampling loop."""
import gzip
import json
import os
import random

import pytest
import torch

from helpers import GOLD, build_hf_llama, load_cases
from oracle import lookahead as LA
from oracle import llama_ref as LR

pytestmark = pytest.mark.gpu
with gzip.open(os.path.join(GOLD, "sample_traces.json.gz"), "rt") as f:
    S_CASES = json.load(f)


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


def test_sampling_topk1_equals_greedy():
    from lookaheaddecoding_b200 import LookaheadEngine
    from lookaheaddecoding_b200.sampling import sample_lookahead
    from test_gpu_e2e import assert_same_or_tie
    c = load_cases()["tiny_bf16_w15n5g15_pool"]
    model, w = build_hf_llama(c["model"], c["weight_seed"])
    eng = LookaheadEngine(model, c["W"], c["N"], c["G"], pool_from_prompt=True, max_total_len=len(c["prompt"]) + 64)
    greedy = eng.generate(c["prompt"], 64, rng=random.Random(2))
    steps_g = eng.last_steps
    torch.manual_seed(0)
    sampled = sample_lookahead(eng, c["prompt"], 64, make_warper(1.0, 1, 1.0), rng=random.Random(2))
    assert len(sampled) == len(greedy)
    assert_same_or_tie(sampled, greedy, c["model"], w, "top_k=1 sampling vs greedy")
    print(f"greedy steps {steps_g}, sampling(top_k=1) steps {eng.last_steps}")
    assert eng.last_steps < 64          # the verification branch accepted something
    eng.close()


@pytest.mark.parametrize("name", ["s_fp32_smallv_t1_w5n3g5", "s_fp32_smallv_eos_w6n4g6", "s_bf16_t08_w15n5g15",
                                  "s_fp32_t07_k50_p09_w7n4g7"])
def test_sampling_seeded_matches_oracle_on_device(name):
    from lookaheaddecoding_b200 import LookaheadEngine
    from lookaheaddecoding_b200.sampling import sample_lookahead
    c = S_CASES[name]
    model, w = build_hf_llama(c["model"], c["weight_seed"])          # bf16 on the GPU (the engine's dtype)
    eos = [c["eos_token_id"]] if c["eos_token_id"] is not None else []
    eng = LookaheadEngine(model, c["W"], c["N"], c["G"], pool_from_prompt=c["pool_from_prompt"],
                          max_total_len=len(c["prompt"]) + c["max_new"])
    warper = make_warper(c["temperature"], c["top_k"], c["top_p"])
    torch.manual_seed(c["seed"] + 1000)
    ours = sample_lookahead(eng, c["prompt"], c["max_new"], warper, eos_token_ids=eos, rng=random.Random(c["seed"]))
    steps = eng.last_steps
    eng.close()
    wb = {k: v.to(torch.bfloat16) for k, v in w.items()}
    om = LR.OracleLlama(c["model"], wb, device="cuda")
    torch.manual_seed(c["seed"] + 1000)
    ref, ref_steps = LA.sample_lookahead(c["prompt"], c["max_new"], c["W"], c["N"], c["G"], om, warper=warper,
                                         pool_from_prompt=c["pool_from_prompt"], eos_token_id=eos or None,
                                         rng=random.Random(c["seed"]))
    n = min(len(ours), len(ref))
    first = next((i for i in range(n) if ours[i] != ref[i]), None)
    print(f"{name}: ours {len(ours) - len(c['prompt'])} tokens / {steps} steps; oracle {len(ref) - len(c['prompt'])} / "
          f"{ref_steps}; first divergence {first}")
    P = len(c["prompt"])
    assert all(0 <= t < c["model"]["vocab"] for t in ours[P:])
    if first is None:
        assert len(ours) == len(ref)
        return
    # a flipped draw is only legitimate where both tokens are live candidates of the warped distribution
    om2 = LR.OracleLlama(c["model"], wb, device="cuda")
    vis = torch.tril(torch.ones(first, first, dtype=torch.bool))
    logits = om2.forward_rows(ours[:first], list(range(first)), vis, 0)[-1:]
    probs = torch.softmax(warper(torch.tensor([ours[:first]], device="cuda"), logits), dim=-1)[0]
    # top-k / top-p cut-offs move with the last bf16 bit of a logit, so truncated distributions may part earlier
    min_match = 8 if (not c["top_k"] and c["top_p"] >= 1.0) else 1
    assert first - P >= min_match, f"diverged after only {first - P} tokens"
    # both tokens must be live candidates of the UNtruncated temperature distribution (top-k / top-p membership itself
    # flips with the last bit of a logit)
    from transformers.generation.logits_process import LogitsProcessorList, TemperatureLogitsWarper
    tw = LogitsProcessorList([w_ for w_ in warper if isinstance(w_, TemperatureLogitsWarper)])
    probs_t = torch.softmax(tw(torch.tensor([ours[:first]], device="cuda"), logits), dim=-1)[0]
    assert probs_t[ours[first]] > 1e-5 and probs_t[ref[first]] > 1e-5, "divergence at a token without probability mass"


def test_generate_do_sample_routes_to_sampling_loop(monkeypatch):
    import lade
    c = load_cases()["tiny_bf16_w5n3g3"]
    model, w = build_hf_llama(c["model"], c["weight_seed"])
    model.generation_config.pad_token_id = 0
    model.generation_config.eos_token_id = None
    ids = torch.tensor([c["prompt"]], device="cuda")
    monkeypatch.setenv("USE_LADE", "1")
    lade.augment_all()
    try:
        lade.config_lade(LEVEL=c["N"], WINDOW_SIZE=c["W"], GUESS_SET_SIZE=c["G"], DEBUG=1, POOL_FROM_PROMPT=True)
        random.seed(1); torch.manual_seed(1)
        a = model.generate(ids, attention_mask=torch.ones_like(ids), max_new_tokens=32, do_sample=True, temperature=0.8,
                           top_k=0, top_p=1.0)
        random.seed(1); torch.manual_seed(1)
        b = model.generate(ids, attention_mask=torch.ones_like(ids), max_new_tokens=32, do_sample=True, temperature=0.8,
                           top_k=0, top_p=1.0)
        assert a.shape == (1, len(c["prompt"]) + 32) and torch.equal(a, b)      # same seeds -> same draws
        random.seed(2); torch.manual_seed(2)
        d = model.generate(ids, attention_mask=torch.ones_like(ids), max_new_tokens=32, do_sample=True, temperature=0.8,
                           top_k=0, top_p=1.0)
        assert not torch.equal(a, d)
        with pytest.raises(Exception):      # warpers outside {temperature, top-k, top-p} are rejected (decoding.py:377)
            model.generate(ids, attention_mask=torch.ones_like(ids), max_new_tokens=4, do_sample=True, repetition_penalty=1.3)
    finally:
        lade.restore_generate()
