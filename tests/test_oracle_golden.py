"""Pin the CPU oracle against fixtures produced by the unmodified reference (tests/golden/gen_golden.py)."""
import gzip
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import lookahead as LA
from oracle import llama_ref as LR

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_cases():
    with gzip.open(os.path.join(GOLD, "greedy_traces.json.gz"), "rt") as f:
        return json.load(f)


def load_masks():
    with gzip.open(os.path.join(GOLD, "masks.json.gz"), "rt") as f:
        return json.load(f)


CASES = load_cases()


def rows_to_bool(rows):
    return np.array([[ch == "1" for ch in r] for r in rows], dtype=bool)


def check_oracle_greedy_trace(c):
    """Replay one reference run (golden trace dict `c`) through the oracle: every step's rows, guesses, argmax
    tokens, KV bookkeeping and mask, then the output ids and the final pool must be identical."""
    dtype = getattr(torch, c["dtype"])
    w = LR.init_weights(c["model"], seed=c["weight_seed"], dtype=dtype)
    om = LR.OracleLlama(c["model"], w)
    trace, pool = [], {}
    ids, steps = LA.greedy_lookahead(
        c["prompt"], c["max_new"], c["W"], c["N"], c["G"], om.step_fn, om.compact_fn,
        pool_from_prompt=c["pool_from_prompt"], eos_token_id=c["eos_token_id"],
        rng=random.Random(c["py_seed"]), trace=trace, token_map_out=pool)
    assert steps == c["n_steps"] == len(trace)
    assert ids == c["output_ids"]
    for i, (t, g) in enumerate(zip(trace, c["steps"])):
        n_in = len(g["input_ids"]) if i == 0 else 1
        flat = g["input_ids"][-n_in:]
        for lvl in g["past_tokens"][: g["fill_level"] + 1]:
            flat = flat + lvl
        flat = flat + (g["guess_tokens"] or [])
        assert t.ids == flat, f"step {i} rows"
        assert t.guess_tokens == g["guess_tokens"], f"step {i} guesses"
        assert t.first_guess == g["first_guess"] and t.inp_tokens == g["inp_tokens"], f"step {i} argmax"
        assert t.guess_results == g["guess_results"], f"step {i} guess argmax"
        assert t.kv_len + n_in == g["kvcache_len"] and t.kv_len + len(t.ids) == g["step_len"]
        if g["mask_rows"] is not None:
            lay = LA.layout_from_shape(t.level_sizes, n_in, len(t.guess_tokens or []), c["N"] - 1,
                                       is_prefill=(i == 0))
            want = rows_to_bool(g["mask_rows"])
            np.testing.assert_array_equal(LA.step_mask(lay), want[:, t.kv_len:], err_msg=f"step {i} mask")
    got_pool = {str(k): [list(t) for t in v] for k, v in pool.items()}
    assert got_pool == c["final_pool"]


def test_mask_predicate_matches_reference_builder():
    """oracle.row_sees == j_make_causal_mask_multilevel (modeling_llama.py:115) on 100+ shapes incl. LP."""
    masks = load_masks()
    assert len(masks) >= 100
    for fx in masks:
        lay = LA.layout_from_shape(fx["level_sizes"], fx["n_extra_input"] + 1, fx["guess_len"], fx["guess_size"])
        got = LA.step_mask(lay)
        want = rows_to_bool(fx["rows"])
        kv = fx["kv"]
        assert want[:, :kv].all(), "cache columns are visible to every row"
        np.testing.assert_array_equal(got, want[:, kv:], err_msg=str({k: fx[k] for k in fx if k != "rows"}))


def test_media_mask_png_kat():
    """The N=4, W=5, 2-guess mask of media/mask.png (SURVEY.md App. B dump)."""
    lay = LA.layout_from_shape([4, 5, 5], 1, 6, 3)
    m = LA.step_mask(lay)
    txt = ["".join("#" if v else "." for v in row) for row in m]
    assert txt[0] == "#" + "." * 20
    assert txt[5] == "#....#" + "." * 15           # L1[0]: block0[0], self
    assert txt[6] == "##....#" + "." * 14          # L1[1]
    assert txt[10] == "#....#....#" + "." * 10     # L2[0]
    assert txt[15] == "#" + "." * 14 + "#....."     # guess0[0]: input, self
    assert txt[17] == "#" + "." * 14 + "###..."
    assert txt[18] == "#" + "." * 17 + "#.."


@pytest.mark.parametrize("name", sorted(CASES))
def test_greedy_trace_matches_reference(name):
    """Step-by-step equality of the restated loop + model with the reference's own run."""
    check_oracle_greedy_trace(CASES[name])


def test_position_ids_match_reference():
    c = CASES["tiny_fp32_w7n5g7"]
    w = LR.init_weights(c["model"], seed=c["weight_seed"], dtype=torch.float32)
    om = LR.OracleLlama(c["model"], w)
    trace = []
    LA.greedy_lookahead(c["prompt"], 24, c["W"], c["N"], c["G"], om.step_fn, om.compact_fn,
                        rng=random.Random(c["py_seed"]), trace=trace)
    # reference: input position ids then window then guesses (modeling_llama.py:1479-1503)
    for t, g in zip(trace, c["steps"]):
        lst = g["position_ids"][-1]
        assert t.pos[: len(g["position_ids"])][-1] == lst
        n_in = len(t.ids) - sum(t.level_sizes) - len(t.guess_tokens or [])
        assert t.pos[n_in:n_in + t.level_sizes[0]] == list(range(lst + 1, lst + 1 + t.level_sizes[0]))
        if t.guess_tokens:
            gs = c["N"] - 1
            assert t.pos[-len(t.guess_tokens):] == list(range(lst + 1, lst + 1 + gs)) * (len(t.guess_tokens) // gs)


def test_lookahead_equals_plain_greedy_fp32():
    """The self-evident invariant the reference claims (minimal.py:55) holds for the restatement."""
    c = CASES["cfg1_fp32_w5n3g3"]
    w = LR.init_weights(c["model"], seed=c["weight_seed"], dtype=torch.float32)
    om = LR.OracleLlama(c["model"], w)
    plain = om.plain_greedy(c["prompt"], 32)
    assert plain == c["output_ids"][: len(plain)]


@pytest.mark.parametrize("name", ["attn_tiny_bf16_w15n5g15_pool", "attn_gqa_bf16_w15n5g15", "attn_tiny_bf16_w5n3g3"])
def test_eager_attention_restatement_bitexact(name):
    """oracle.llama_ref.eager_attention reproduces the reference module's output bit-for-bit (CPU bf16)."""
    fx = torch.load(os.path.join(GOLD, name + ".pt"))
    vis = torch.from_numpy(rows_to_bool(fx["mask_rows"]))
    kv_len = fx["kv_len"]
    mask = LR.additive_mask(vis[:, kv_len:], kv_len, fx["q"].dtype)
    o = LR.eager_attention(fx["q"], fx["k"], fx["v"], mask, fx["q"].shape[0] // fx["k"].shape[0])
    o = o.transpose(0, 1).reshape(fx["o"].shape)
    assert torch.equal(o, fx["o"])
