"""Edge-case configurations of the greedy lookahead loop: oracle vs traces of the unmodified reference.

tests/golden/greedy_edge_traces.json.gz (tests/golden/gen_golden_edge.py): one-token prompt, prompt shorter than N,
one / two new tokens, W=1, W=2 with G=1, G=1 under a wide window, N=3 with a wide window, EOS on the first generated
token.  CPU-only for now: these shapes have not been replayed through the device state machine yet (round-2 item)."""
import gzip
import json
import os

import pytest

from test_oracle_golden import GOLD, check_oracle_greedy_trace


def _load():
    with gzip.open(os.path.join(GOLD, "greedy_edge_traces.json.gz"), "rt") as f:
        return json.load(f)


EDGE = _load()


@pytest.mark.parametrize("name", sorted(EDGE["cases"]))
def test_edge_case_trace_matches_reference(name):
    check_oracle_greedy_trace(EDGE["cases"][name])


def test_edge_fixture_covers_the_intended_shapes():
    cases = EDGE["cases"]
    assert len(cases[ "edge_p1_w5n3g3_pool"]["prompt"]) == 1
    assert len(cases["edge_p2_w7n5g7_pool"]["prompt"]) < cases["edge_p2_w7n5g7_pool"]["N"]
    assert cases["edge_new1_w5n3g3"]["n_steps"] == 1 and cases["edge_new2_w5n3g3_pool"]["n_steps"] == 2
    assert cases["edge_w1n3g1_pool"]["W"] == 1 and cases["edge_g1_w15n5_pool"]["G"] == 1
    eos = cases["edge_eos_first_w5n3g3"]
    assert eos["n_generated"] == 1 and eos["output_ids"][-1] == eos["eos_token_id"]
    # every multi-step case accepted at least one guessed token somewhere (the verification branch ran)
    for name in ("edge_p1_w5n3g3_pool", "edge_p2_w7n5g7_pool", "edge_w1n3g1_pool", "edge_w2n4g1_pool",
                 "edge_g1_w15n5_pool", "edge_n3_w20g20_pool"):
        assert cases[name]["n_steps"] < cases[name]["n_generated"], name
    # GUESS_SET_SIZE=0 trips an assert inside the reference (lade/decoding.py:48); this repo treats G=0 as
    # "verification off" (plain greedy), which the reference cannot express
    assert "edge_g0_w5n4" in EDGE["reference_rejects"]


def _all_traces():
    from helpers import load_cases
    out = dict(load_cases())
    out.update(EDGE["cases"])
    return out


@pytest.mark.parametrize("name", sorted(_all_traces()))
def test_host_row_bound_covers_every_reference_step(name):
    """lade_step_rows_bound (host-only C-ABI call, sizes q_pad and the captured graph) must be an upper bound of the
    rows the reference forwards at every step, and exact for the prefill step and for steps with a full guess set."""
    import ctypes as C
    from helpers import make_lade_config
    from lookaheaddecoding_b200 import _cabi
    lib = _cabi.load()
    c = _all_traces()[name]
    W, N, G, P = c["W"], c["N"], c["G"], len(c["prompt"])
    cfg = make_lade_config(W, N, G, c["model"]["vocab"], P + c["max_new"] + N + 8, pool=c["pool_from_prompt"])
    for i, g in enumerate(c["steps"]):
        n_in = len(g["input_ids"]) if i == 0 else 1
        rows = n_in + sum(len(l) for l in g["past_tokens"][: g["fill_level"] + 1]) + len(g["guess_tokens"] or [])
        bound = lib.lade_step_rows_bound(C.byref(cfg), P, i)
        assert bound >= rows, f"step {i}: bound {bound} < {rows} rows"
        if i == 0:
            assert bound == rows
        if g["guess_tokens"] is not None and len(g["guess_tokens"]) == G * (N - 1) and g["fill_level"] == N - 2:
            assert bound == rows, f"steady step {i}"
