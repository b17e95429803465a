"""n-gram pool (LRU, G tuples per key): oracle and device vs known-answer vectors of the reference's own functions.

tests/golden/pool_fuzz.json.gz was produced by tests/golden/gen_golden_pool.py running the unmodified
update_token_map / append_new_generated_pool / fill_pool_with_prompt (lade/decoding.py:37,80,104) on tiny
vocabularies, so that keys collide, tuples repeat (move-to-back) and full keys evict their oldest tuple."""
import ctypes as C
import gzip
import json
import os

import numpy as np
import pytest

from helpers import GOLD, make_lade_config
from oracle import lookahead as LA


def load_pool_cases():
    with gzip.open(os.path.join(GOLD, "pool_fuzz.json.gz"), "rt") as f:
        return json.load(f)


POOL_CASES = load_pool_cases()
IDS = [f"N{c['N']}W{c['W']}G{c['G']}V{c['vocab']}" for c in POOL_CASES]


def _as_pool(d):
    return {int(k): [tuple(t) for t in v] for k, v in d.items()}


@pytest.mark.parametrize("case", POOL_CASES, ids=IDS)
def test_oracle_pool_matches_reference_vectors(case):
    N, W, G = case["N"], case["W"], case["G"]
    token_map = {}
    evictions = moves = 0
    for i, op in enumerate(case["ops"]):
        before = {k: list(v) for k, v in token_map.items()}
        if op["op"] == "fill":
            LA.fill_pool_with_prompt(op["prompt"], token_map, N, G)
        elif op["op"] == "update":
            LA.update_token_map(token_map, op["lst"], op["past"], op["new"], N, W, G)
        else:
            LA.append_new_generated_pool(op["tokens"], token_map, N, G)
        assert {k: v for k, v in token_map.items()} == _as_pool(op["pool"]), f"op {i} ({op['op']})"
        for k, v in token_map.items():
            assert len(v) <= G and len(set(v)) == len(v)
            old = before.get(k, [])
            evictions += sum(1 for t in old if t not in v)
            moves += int(bool(old) and old != v[: len(old)] and set(old) <= set(v))
    # the vectors are only worth something if the interesting paths were taken somewhere
    if case["vocab"] <= 7 and G <= 3:
        assert evictions > 0


def test_pool_vectors_cover_lru_paths():
    """Across the fixture: evictions, move-to-back re-insertions and ignored wrong-length appends all occur."""
    evict = move = ignored = 0
    for c in POOL_CASES:
        prev = {}
        for op in c["ops"]:
            cur = _as_pool(op["pool"])
            if op["op"] == "append" and len(op["tokens"]) != c["N"]:
                assert cur == prev
                ignored += 1
            for k, v in cur.items():
                old = prev.get(k, [])
                evict += sum(1 for t in old if t not in v)
                if old and set(old) <= set(v) and old != v[: len(old)]:
                    move += 1
            prev = cur
    assert evict > 50 and move >= 5 and ignored > 3, (evict, move, ignored)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in POOL_CASES if len(c["ops"][0]["prompt"]) >= 1],
                         ids=[i for i, c in zip(IDS, POOL_CASES) if len(c["ops"][0]["prompt"]) >= 1])
def test_device_prompt_pool_fill_matches_reference_vectors(case):
    """lade_ctx_reset with POOL_FROM_PROMPT builds the pool on the device (fill_pool_from_prompt_kernel)."""
    import torch
    from lookaheaddecoding_b200 import _cabi
    from lookaheaddecoding_b200._cabi import check

    lib = _cabi.load()
    N, W, G, vocab = case["N"], case["W"], case["G"], case["vocab"]
    prompt = case["ops"][0]["prompt"]
    P, GS, WCAP = len(prompt), N - 1, W + N - 3
    cfg = make_lade_config(W, N, G, vocab, P + 16, pool=True)
    ctx = C.c_void_p()
    check(lib.lade_ctx_create(C.byref(cfg), C.byref(ctx)), "create")
    stream = torch.cuda.current_stream().cuda_stream
    pr = np.asarray(prompt, dtype=np.int32)
    w0 = np.asarray([prompt[0]] * WCAP, dtype=np.int32)
    check(lib.lade_ctx_reset(ctx, stream, pr.ctypes.data, P, w0.ctypes.data, WCAP, P + 8), "reset")
    torch.cuda.synchronize()
    cnt = np.zeros(vocab, dtype=np.int32)
    tup = np.zeros((vocab, max(G, 1), GS), dtype=np.int32)
    check(lib.lade_ctx_pool_snapshot(ctx, stream, cnt.ctypes.data, tup.ctypes.data), "snapshot")
    got = {int(k): [tuple(x) for x in tup[k, : cnt[k]].tolist()] for k in np.nonzero(cnt)[0]}
    want = {k: v for k, v in _as_pool(case["ops"][0]["pool"]).items() if v}
    lib.lade_ctx_destroy(ctx)
    assert got == want
