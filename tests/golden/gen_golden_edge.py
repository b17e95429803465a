#!/usr/bin/env python
"""Generates tests/golden/greedy_edge_traces.json.gz: per-step traces of the UNMODIFIED reference greedy loop
(/root/reference/lade/decoding.py:697) on edge-case configurations -- one-token and shorter-than-N prompts, one or
two new tokens, minimal windows, G=1 pools, EOS on the first token.  CPU-only fixture for tests/test_oracle_edge.py
(same trace format as gen_golden.py's greedy cases)."""
import gzip
import json
import os
import sys
import traceback
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

from oracle import ref_shim as R  # noqa: E402
from gen_golden import TINY, trace_reference_greedy  # noqa: E402

EDGE_CASES = [
    # name, W, N, G, pool_from_prompt, prompt_len, max_new, weight seed, prompt seed, eos ("first" = first new token)
    ("edge_p1_w5n3g3_pool", 5, 3, 3, True, 1, 24, 0, 11, None),
    ("edge_p2_w7n5g7_pool", 7, 5, 7, True, 2, 24, 0, 12, None),
    ("edge_new1_w5n3g3", 5, 3, 3, False, 12, 1, 0, 13, None),
    ("edge_new2_w5n3g3_pool", 5, 3, 3, True, 12, 2, 0, 14, None),
    ("edge_w1n3g1_pool", 1, 3, 1, True, 16, 24, 0, 15, None),
    ("edge_w2n4g1_pool", 2, 4, 1, True, 16, 32, 0, 16, None),
    ("edge_g1_w15n5_pool", 15, 5, 1, True, 32, 48, 0, 17, None),
    ("edge_n3_w20g20_pool", 20, 3, 20, True, 24, 48, 0, 18, None),
    ("edge_eos_first_w5n3g3", 5, 3, 3, True, 12, 16, 0, 19, "first"),
    ("edge_g0_w5n4", 5, 4, 0, False, 12, 16, 0, 20, None),
]


def main():
    cases, skipped = {}, {}
    for (name, W, N, G, pool, P, max_new, wseed, pseed, eos) in EDGE_CASES:
        cfg = R.make_llama_config(**TINY)
        model = R.build_reference_model(cfg, seed=wseed, dtype=torch.float32)
        torch.manual_seed(pseed)
        prompt = torch.randint(3, TINY["vocab"], (1, P))
        lade_cfg = dict(WINDOW_SIZE=W, LEVEL=N, GUESS_SET_SIZE=G, DEBUG=1, POOL_FROM_PROMPT=int(pool))
        try:
            eos_id = None
            if eos == "first":
                ids0, *_ = trace_reference_greedy(model, prompt, max_new, lade_cfg, py_seed=pseed)
                eos_id = ids0[P]
            ids, steps, pool_d, log, _ = trace_reference_greedy(model, prompt, max_new, lade_cfg, py_seed=pseed, eos=eos_id)
        except Exception as e:   # the reference itself rejects / crashes on this configuration: record, do not pin
            skipped[name] = f"{type(e).__name__}: {e}"[:200]
            print(f"{name}: reference failed: {skipped[name]}")
            traceback.print_exc(limit=2)
            continue
        cases[name] = dict(model=TINY, dtype="float32", W=W, N=N, G=G, pool_from_prompt=bool(pool), weight_seed=wseed,
                           prompt=prompt[0].tolist(), max_new=max_new, py_seed=pseed, eos_token_id=eos_id, output_ids=ids,
                           n_steps=log[1], n_generated=log[0], steps=steps, final_pool=pool_d)
        print(f"{name}: generated {log[0]} tokens in {log[1]} steps")
    with gzip.open(os.path.join(HERE, "greedy_edge_traces.json.gz"), "wt") as f:
        json.dump(dict(cases=cases, reference_rejects=skipped), f, separators=(",", ":"))


if __name__ == "__main__":
    main()
