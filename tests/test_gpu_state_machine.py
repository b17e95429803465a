"""Bit-exact parity of the device state machine (lade_step_layout / lade_accept_update / pool / window)
against the reference's own per-step traces (tests/golden, produced by the unmodified reference).

The float side is taken out of the loop: the argmax tokens the reference's model produced at each step
are fed to the accept kernel, so every integer result must be identical."""
import ctypes as C

import numpy as np
import pytest
import torch

from helpers import load_cases, make_lade_config, rows_to_bool

pytestmark = pytest.mark.gpu


def _all_cases():
    """The 10 regular traces + the 9 edge-case traces (1-token prompt, prompt < N, 1-2 new tokens, W=1, G=1, N=3 under a
    wide window, EOS on the first token; tests/golden/gen_golden_edge.py)."""
    import gzip
    import json
    import os
    from helpers import GOLD
    out = dict(load_cases())
    with gzip.open(os.path.join(GOLD, "greedy_edge_traces.json.gz"), "rt") as f:
        out.update(json.load(f)["cases"])
    return out


CASES = _all_cases()


def _decode_rowdesc(rd):
    rd = rd.astype(np.int64) & 0xFFFFFFFF
    return (rd >> 30) & 3, (rd >> 15) & 0x7FFF, rd & 0x7FFF


def _mask_from_rowdesc(rd, q_len, level_offset):
    cls, blk, idx = _decode_rowdesc(rd)
    m = np.zeros((q_len, q_len), dtype=bool)
    for r in range(q_len):
        for c in range(q_len):
            if cls[r] == 0:
                m[r, c] = cls[c] == 0 and c <= r
            elif cls[r] == 1:
                if cls[c] == 0:
                    m[r, c] = True
                elif cls[c] == 1:
                    m[r, c] = (idx[c] <= idx[r]) if blk[c] == 0 else (blk[c] <= blk[r] and idx[c] == idx[r])
            elif cls[r] == 2:
                m[r, c] = c <= level_offset or (cls[c] == 2 and blk[c] == blk[r] and idx[c] <= idx[r])
            else:
                m[r, c] = r == c
    return m


@pytest.mark.parametrize("name", sorted(CASES))
def test_device_state_machine_matches_reference_trace(name):
    from lookaheaddecoding_b200 import _cabi
    from lookaheaddecoding_b200._cabi import check

    lib = _cabi.load()
    c = CASES[name]
    W, N, G = c["W"], c["N"], c["G"]
    GS, WCAP = N - 1, W + N - 3
    V = c["model"]["vocab"]
    P = len(c["prompt"])
    max_length = P + c["max_new"]
    eos = [c["eos_token_id"]] if c["eos_token_id"] is not None else []
    cfg = make_lade_config(W, N, G, V, max_length + N + 8, pool=c["pool_from_prompt"], eos=eos)
    ctx = C.c_void_p()
    check(lib.lade_ctx_create(C.byref(cfg), C.byref(ctx)), "create")
    dev = torch.device("cuda")
    stream = torch.cuda.current_stream().cuda_stream
    steps = c["steps"]
    prompt = np.asarray(c["prompt"], dtype=np.int32)
    win0 = np.asarray(steps[0]["past_tokens"][0], dtype=np.int32)
    assert len(win0) == WCAP
    check(lib.lade_ctx_reset(ctx, stream, prompt.ctypes.data, P, win0.ctypes.data, WCAP, max_length), "reset")
    torch.cuda.synchronize()
    lm_cap = 1 + WCAP + G * GS
    q_cap = max(P + WCAP, GS * (W + G)) + 5
    i32 = dict(dtype=torch.int32, device=dev)
    ids, pos, rd = torch.zeros(q_cap, **i32), torch.zeros(q_cap, **i32), torch.zeros(q_cap, **i32)
    lm_rows, meta = torch.zeros(lm_cap, **i32), torch.zeros(_cabi.META_INTS, **i32)
    res = torch.zeros(_cabi.RES_INTS, **i32)
    out_ids = list(c["prompt"])
    for i, g in enumerate(steps):
        n_in = P if i == 0 else 1
        flat = g["input_ids"][-n_in:]
        flat_pos = g["position_ids"][-n_in:]
        lst = g["position_ids"][-1]
        for ll, lvl in enumerate(g["past_tokens"][: g["fill_level"] + 1]):
            flat = flat + lvl
            if ll == 0:
                flat_pos = flat_pos + list(range(lst + 1, lst + 1 + len(lvl)))
            else:
                off = len(g["past_tokens"][0]) + 1 - len(lvl)
                flat_pos = flat_pos + list(range(lst + ll + off, lst + ll + off + len(lvl)))
        gt = g["guess_tokens"] or []
        flat = flat + gt
        flat_pos = flat_pos + list(range(lst + 1, lst + 1 + GS)) * (len(gt) // GS)
        q_len = len(flat)
        q_pad = q_len + 3          # also exercise PAD rows
        mw = (q_pad + 31) // 32 + 1
        rowmask = torch.zeros(q_pad * mw, **i32)
        check(lib.lade_step_layout(ctx, stream, q_pad, ids.data_ptr(), pos.data_ptr(), rd.data_ptr(),
                                   lm_rows.data_ptr(), meta.data_ptr(), rowmask.data_ptr(), mw), "layout")
        m = meta.cpu().numpy()
        assert m[_cabi.M_Q_LEN] == q_len, f"step {i}"
        assert m[_cabi.M_KV_LEN] + n_in == g["kvcache_len"] and m[_cabi.M_KV_LEN] + q_len == g["step_len"]
        assert ids[:q_len].cpu().tolist() == flat, f"step {i} ids"
        assert pos[:q_len].cpu().tolist() == flat_pos, f"step {i} pos"
        rd_h = rd[:q_pad].cpu().numpy()
        assert (((rd_h[q_len:].astype(np.int64) & 0xFFFFFFFF) >> 30) == 3).all()
        if g["mask_rows"] is not None:
            want = rows_to_bool(g["mask_rows"])[:, m[_cabi.M_KV_LEN]:]
            got = _mask_from_rowdesc(rd_h, q_len, int(m[_cabi.M_LEVEL_OFFSET]))
            np.testing.assert_array_equal(got, want, err_msg=f"step {i} mask")
            if i > 0:   # the bitmask the attention kernels consume (prefill steps carry none: plain causal)
                wd = rowmask.cpu().numpy().view(np.uint32).reshape(q_pad, mw)
                bits = np.unpackbits(wd.view(np.uint8).reshape(q_pad, mw * 4), axis=-1, bitorder="little").astype(bool)
                np.testing.assert_array_equal(bits[:q_len, :q_len], want, err_msg=f"step {i} rowmask")
                assert not bits[q_len:].any() and not bits[:, q_len:].any()
        tiny = int(m[_cabi.M_TINY])
        lmr = lm_rows.cpu().numpy()
        assert lmr[0] == n_in - 1
        assert lmr[1:1 + tiny].tolist() == list(range(q_len - len(gt) - tiny, q_len - len(gt)))
        assert lmr[1 + WCAP:1 + WCAP + len(gt)].tolist() == list(range(q_len - len(gt), q_len))
        # feed the reference model's argmax tokens
        am = np.zeros(lm_cap, dtype=np.int32)
        am[0] = g["first_guess"]
        am[1:1 + len(g["inp_tokens"])] = g["inp_tokens"]
        am[1 + WCAP:1 + WCAP + len(g["guess_results"])] = g["guess_results"]
        am_d = torch.from_numpy(am).to(dev)
        check(lib.lade_accept_update(ctx, stream, am_d.data_ptr(), meta.data_ptr(), res.data_ptr()), "accept")
        r = res.cpu().numpy()
        n_emit = int(r[_cabi.R_N_EMIT])
        out_ids += r[_cabi.R_HITS:_cabi.R_HITS + n_emit].tolist()
        if i + 1 < len(steps):
            nxt = steps[i + 1]
            assert out_ids == nxt["input_ids"][: len(out_ids)] or out_ids[-1:] == nxt["input_ids"][-1:], f"step {i}"
            assert int(r[_cabi.R_KV_LEN]) + 1 == nxt["kvcache_len"], f"step {i} kv_len"
            assert not r[_cabi.R_DONE]
        else:
            assert r[_cabi.R_DONE]
    assert out_ids[:max_length] == c["output_ids"]
    assert int(r[_cabi.R_STEPS]) == c["n_steps"]
    # pool snapshot == the reference's token_map
    cnt = np.zeros(V, dtype=np.int32)
    tup = np.zeros((V, G, GS), dtype=np.int32)
    check(lib.lade_ctx_pool_snapshot(ctx, stream, cnt.ctypes.data, tup.ctypes.data), "snapshot")
    got = {str(k): tup[k, : cnt[k]].tolist() for k in np.nonzero(cnt)[0]}
    assert got == c["final_pool"]
    lib.lade_ctx_destroy(ctx)


def test_argmax_rows_ties_and_unaligned():
    from lookaheaddecoding_b200 import _cabi
    from lookaheaddecoding_b200._cabi import check
    lib = _cabi.load()
    torch.manual_seed(0)
    for vocab, ld in [(32000, 32000), (32016, 32016), (1000, 1003), (7, 8)]:
        x = (torch.randn(37, ld, device="cuda") * 4).round().to(torch.bfloat16)   # many exact ties
        out = torch.zeros(37, dtype=torch.int32, device="cuda")
        check(lib.lade_argmax_rows(torch.cuda.current_stream().cuda_stream, x.data_ptr(), 37, vocab, ld, out.data_ptr()))
        want = torch.argmax(x[:, :vocab].float().cpu(), dim=-1)      # CPU argmax: lowest index on ties
        assert out.cpu().tolist() == want.tolist()
