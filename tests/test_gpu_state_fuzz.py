"""Fuzz parity of the device state machine against the oracle on synthetic token streams.

A deterministic fake "model" (next token = f(token id, position) over a SMALL vocabulary) makes n-gram hits,
full-length accepts, LRU evictions, pool re-orderings and EOS stops frequent -- far more of them than the
random-init goldens produce.  The oracle loop (oracle.lookahead.greedy_lookahead, pinned to the reference) and
the CUDA state machine (lade_step_layout / lade_accept_update) are driven with the same fake model; every
step's rows, position ids, hits and the final pool / ids must be identical."""
import ctypes as C
import random

import numpy as np
import pytest
import torch

from helpers import make_lade_config
from oracle import lookahead as LA

pytestmark = pytest.mark.gpu


def fake_token(tok, pos, vocab, salt):
    return int((tok * 7 + pos * 3 + salt + (tok * pos) % 5) % vocab)


@pytest.mark.parametrize("W,N,G,vocab,pool,eos,P,max_new,salt", [
    (5, 3, 3, 11, True, None, 12, 120, 0),
    (15, 5, 15, 23, True, None, 40, 200, 1),
    (15, 5, 15, 23, False, None, 40, 200, 2),
    (7, 4, 2, 31, True, 1, 16, 150, 3),        # EOS inside an accepted n-gram (decoding.py:1168-1173)
    (7, 4, 2, 41, True, 9, 16, 150, 3),
    (20, 7, 20, 31, True, None, 64, 160, 4),
    (3, 8, 4, 5, True, None, 10, 100, 5),
    (60, 8, 60, 37, False, None, 30, 80, 6),       # the reference's defaults (decoding.py:854-857)
    (1, 3, 1, 5, True, None, 8, 60, 7),
])
def test_device_state_machine_fuzz(W, N, G, vocab, pool, eos, P, max_new, salt):
    from lookaheaddecoding_b200 import _cabi
    from lookaheaddecoding_b200._cabi import check

    lib = _cabi.load()
    rnd = random.Random(salt)
    prompt = [rnd.randrange(vocab) for _ in range(P)]
    GS, WCAP = N - 1, W + N - 3

    def step_fn(lay, kv_len):
        outs = [fake_token(t, p, vocab, salt) for t, p in zip(lay.ids, lay.pos)]
        q, lg, win = lay.q_len, lay.n_guess_tok, lay.level_sizes[-1]
        return outs[lay.n_input - 1], outs[q - lg - win:q - lg], outs[q - lg:]

    trace, pool_ref = [], {}
    ref_ids, ref_steps = LA.greedy_lookahead(prompt, max_new, W, N, G, step_fn, lambda *a: None, pool_from_prompt=pool,
                                             eos_token_id=eos, rng=random.Random(salt + 100), trace=trace,
                                             token_map_out=pool_ref)
    window0 = [random.Random(salt + 100).choice(prompt) for _ in range(1)]  # placeholder, replaced below
    r2 = random.Random(salt + 100)
    window0 = [r2.choice(prompt) for _ in range(WCAP)]          # the same draws the oracle made (decoding.py:902)

    max_length = P + max_new
    cfg = make_lade_config(W, N, G, vocab, max_length + N + 8, pool=pool, eos=[eos] if eos is not None else [])
    ctx = C.c_void_p()
    check(lib.lade_ctx_create(C.byref(cfg), C.byref(ctx)), "create")
    stream = torch.cuda.current_stream().cuda_stream
    pr = np.asarray(prompt, dtype=np.int32)
    w0 = np.asarray(window0, dtype=np.int32)
    check(lib.lade_ctx_reset(ctx, stream, pr.ctypes.data, P, w0.ctypes.data, WCAP, max_length), "reset")
    torch.cuda.synchronize()
    lm_cap = 1 + WCAP + G * GS
    q_cap = max(P + WCAP, GS * (W + G)) + 8
    i32 = dict(dtype=torch.int32, device="cuda")
    ids, pos, rd = torch.zeros(q_cap, **i32), torch.zeros(q_cap, **i32), torch.zeros(q_cap, **i32)
    lm_rows, meta, res = torch.zeros(lm_cap, **i32), torch.zeros(_cabi.META_INTS, **i32), torch.zeros(_cabi.RES_INTS, **i32)
    out_ids = list(prompt)
    n_hits = 0
    for i, t in enumerate(trace):
        q_pad = lib.lade_step_rows_bound(C.byref(cfg), P, i)
        check(lib.lade_step_layout(ctx, stream, q_pad, ids.data_ptr(), pos.data_ptr(), rd.data_ptr(), lm_rows.data_ptr(),
                                   meta.data_ptr(), 0, 0), "layout")
        m = meta.cpu().numpy()
        q_len = int(m[_cabi.M_Q_LEN])
        assert q_len == len(t.ids) <= q_pad, f"step {i}"
        ids_h, pos_h = ids[:q_len].cpu().tolist(), pos[:q_len].cpu().tolist()
        assert ids_h == t.ids and pos_h == t.pos, f"step {i} rows"
        outs = [fake_token(a, b, vocab, salt) for a, b in zip(ids_h, pos_h)]
        lmr = lm_rows.cpu().numpy()
        am = np.asarray([outs[r] for r in lmr], dtype=np.int32)       # what lm_head + argmax would deliver
        am_d = torch.from_numpy(am).cuda()
        check(lib.lade_accept_update(ctx, stream, am_d.data_ptr(), meta.data_ptr(), res.data_ptr()), "accept")
        r = res.cpu().numpy()
        assert int(r[_cabi.R_MAX_HIT]) == t.max_hit, f"step {i} max_hit"
        n_emit = int(r[_cabi.R_N_EMIT])
        assert r[_cabi.R_HITS:_cabi.R_HITS + t.max_hit + 1].tolist() == t.hits[: t.max_hit + 1], f"step {i} hits"
        assert int(r[_cabi.R_KV_SRC]) == t.kv_src, f"step {i} kv_src"
        out_ids += r[_cabi.R_HITS:_cabi.R_HITS + n_emit].tolist()
        n_hits += t.max_hit
        assert bool(r[_cabi.R_DONE]) == (i == len(trace) - 1)
    assert out_ids[:max_length] == ref_ids
    assert int(r[_cabi.R_STEPS]) == ref_steps
    cnt = np.zeros(vocab, dtype=np.int32)
    tup = np.zeros((vocab, max(G, 1), GS), dtype=np.int32)
    check(lib.lade_ctx_pool_snapshot(ctx, stream, cnt.ctypes.data, tup.ctypes.data), "snapshot")
    got = {int(k): [tuple(x) for x in tup[k, : cnt[k]].tolist()] for k in np.nonzero(cnt)[0]}
    want = {int(k): [tuple(x) for x in v] for k, v in pool_ref.items() if len(v)}
    assert got == want
    lib.lade_ctx_destroy(ctx)
    print(f"W{W} N{N} G{G} V{vocab}: {len(trace)} steps, {n_hits} accepted guess tokens, {len(ref_ids) - P} generated")
    assert n_hits > 0 or G <= 1
