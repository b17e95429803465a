#!/usr/bin/env python
"""Micro-benchmark of lade_attn_fwd alone (CUDA events, back-to-back over L distinct caches > L2)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lookaheaddecoding_b200 import _cabi  # noqa: E402


def steady_rowmask(W, N, G):
    """Visibility bits of the steady single-GPU lookahead step."""
    GS = N - 1
    q = GS * (W + G)
    vis = np.zeros((q, q), dtype=bool)
    for r in range(q):
        if r < GS * W:
            lvl, j = divmod(r, W)
            vis[r, : j + 1] = True
            for l2 in range(1, lvl + 1):
                vis[r, l2 * W + j] = True
        else:
            e, u = divmod(r - GS * W, GS)
            vis[r, 0] = True
            vis[r, GS * W + e * GS: GS * W + e * GS + u + 1] = True
    mw = (q + 31) // 32 + 1
    bits = np.zeros((q, mw * 32), dtype=bool)
    bits[:, :q] = vis
    words = np.packbits(bits.reshape(q, mw, 32), axis=-1, bitorder="little").view(np.uint32).reshape(q, mw)
    return words.view(np.int32).copy(), mw, q


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kv", type=int, nargs="+", default=[1024, 3072])
    ap.add_argument("--splits", type=int, nargs="+", default=[1, 2, 4])
    ap.add_argument("--impl", type=int, nargs="+", default=[2])
    ap.add_argument("--heads", type=int, default=32)
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--W", type=int, default=15)
    ap.add_argument("--N", type=int, default=5)
    ap.add_argument("--G", type=int, default=15)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    lib = _cabi.load()
    dev = "cuda"
    H, D, L = a.heads, 128, a.layers
    rm_np, mw, q_len = steady_rowmask(a.W, a.N, a.G)
    peak = 6573.8
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        peak = json.load(open(p))["hbm_gbs"]
    for kv in a.kv:
        cap = kv + q_len + 64
        kvc = torch.randn(L, 2, H, cap, D, device=dev, dtype=torch.bfloat16)
        q = torch.randn(H, q_len, D, device=dev, dtype=torch.bfloat16)
        out = torch.empty(q_len, H * D, device=dev, dtype=torch.bfloat16)
        rd = torch.from_numpy(rm_np).to(dev)
        meta = torch.zeros(_cabi.META_INTS, dtype=torch.int32, device=dev)
        for k, v in {_cabi.M_Q_LEN: q_len, _cabi.M_KV_LEN: kv, _cabi.M_N_INPUT: 1, _cabi.M_TINY: a.W,
                     _cabi.M_N_LEVELS: a.N - 1, _cabi.M_N_GUESS_TOK: a.G * (a.N - 1), _cabi.M_PHASE: 2,
                     _cabi.M_Q_PAD: q_len}.items():
            meta[k] = v
        bytes_alg = 2 * kv * H * D * 2 + q_len * H * D * 2 + 2 * q_len * H * D * 2 + q_len * H * D * 2
        for impl in a.impl:
            for ns in a.splits:
                nb = lib.lade_attn_scratch_bytes(q_len, H, D, ns)
                scratch = torch.zeros(nb, dtype=torch.uint8, device=dev)
                def one_pass():
                    st = torch.cuda.current_stream()
                    for l in range(L):
                        _cabi.check(lib.lade_attn_fwd(st.cuda_stream, q.data_ptr(), kvc[l, 0].data_ptr(), kvc[l, 1].data_ptr(),
                                                      out.data_ptr(), rd.data_ptr(), mw, meta.data_ptr(), scratch.data_ptr(), q_len,
                                                      H, H, D, cap, kv + q_len, ns, impl))
                for _ in range(2):
                    one_pass()
                torch.cuda.synchronize()
                # capture one pass over the L caches in a CUDA graph: python/ctypes launch overhead (~20 us per
                # call) would otherwise bound the measurement
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    with torch.cuda.graph(g, stream=side):
                        st = torch.cuda.current_stream()
                        one_pass()
                torch.cuda.current_stream().wait_stream(side)
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(a.reps):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / (a.reps * L)
                gbs = bytes_alg / (us * 1e-6) / 1e9
                print(json.dumps({"kv": kv, "q": q_len, "impl": impl, "splits": ns, "us": round(us, 2),
                                  "GBps": round(gbs, 1), "frac": round(gbs / peak, 4), "alg_MB": round(bytes_alg / 1e6, 2)}))


if __name__ == "__main__":
    main()
