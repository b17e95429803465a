#!/usr/bin/env python
"""Generates tests/golden/pool_fuzz.json.gz: known-answer vectors for the n-gram pool (LRU, G per key).

Runs the UNMODIFIED reference functions `update_token_map`, `append_new_generated_pool` and
`fill_pool_with_prompt` (/root/reference/lade/decoding.py:37,80,104) on seeded random token streams drawn from
tiny vocabularies (so that keys collide, tuples repeat and the LRU evicts), and records the inputs and the resulting
pool after every call.  tests/test_oracle_pool.py replays the inputs through oracle/lookahead.py -- and, on a GPU
box, tests/test_gpu_state_fuzz.py-style device kernels can be checked against the same vectors.
"""
import gzip
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def dump_pool(token_map):
    return {str(k): [list(t) for t in v] for k, v in sorted(token_map.items())}


def main():
    from oracle import ref_shim as R
    decoding, _ = R.load_reference()
    cases = []
    rnd = random.Random(20240917)
    configs = [(3, 5, 3), (4, 7, 2), (5, 15, 15), (5, 15, 1), (7, 20, 20), (6, 4, 5), (3, 1, 1), (5, 2, 7)]
    for (N, W, G) in configs:
        for vocab in (3, 7, 40):
            token_map = {}
            ops = []
            prompt = [rnd.randrange(vocab) for _ in range(rnd.randrange(0, 3 * N + 4))]
            decoding.fill_pool_with_prompt(prompt, token_map, N, G)
            ops.append(dict(op="fill", prompt=prompt, pool=dump_pool(token_map)))
            for _ in range(14):
                kind = rnd.random()
                if kind < 0.6:
                    lst = rnd.randrange(vocab)
                    past = [[rnd.randrange(vocab) for _ in range(W + N - 2 - lv)] for lv in range(N - 1)]
                    # the reference indexes past_tokens[ll][i] for i < W only
                    new = [rnd.randrange(vocab) for _ in range(W)]
                    decoding.update_token_map(token_map, lst, past, new, N, W, G)
                    ops.append(dict(op="update", lst=lst, past=past, new=new, pool=dump_pool(token_map)))
                else:
                    n_tok = N if kind < 0.9 else rnd.choice([N - 1, N + 1])     # wrong lengths are ignored (:81-82)
                    toks = [rnd.randrange(vocab) for _ in range(n_tok)]
                    decoding.append_new_generated_pool(toks, token_map, N, G)
                    ops.append(dict(op="append", tokens=toks, pool=dump_pool(token_map)))
            cases.append(dict(N=N, W=W, G=G, vocab=vocab, ops=ops))
    with gzip.open(os.path.join(HERE, "pool_fuzz.json.gz"), "wt") as f:
        json.dump(cases, f, separators=(",", ":"))
    print(f"{len(cases)} cases, {sum(len(c['ops']) for c in cases)} pool states")


if __name__ == "__main__":
    main()
