#!/usr/bin/env python
"""Micro-benchmark of lade_gemm_bf16 against torch.mm (cuBLAS) on the lookahead-step projection shapes.

Each shape streams `layers` distinct weight matrices (>> L2) back to back from a CUDA graph; the time per launch
is CUDA-event time / layers.  Bytes = weight bytes (N*K*2) + A + C, reported against MEASURED_PEAKS hbm_gbs.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lookaheaddecoding_b200 import _cabi  # noqa: E402

SHAPES_7B = {"qkv": (12288, 4096), "o": (4096, 4096), "gate_up": (22016, 4096), "down": (4096, 11008),
             "lm_head": (32000, 4096)}


def time_graph(fn, reps):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn(s)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn(s)
    torch.cuda.synchronize()
    for _ in range(3):
        g.replay()
    best = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    best.sort()
    return best[len(best) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=120)
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--shapes", nargs="+", default=list(SHAPES_7B))
    ap.add_argument("--configs", nargs="*", default=[], help="extra tile_n:split_k[:depth_cap[:no_prefill]] configs to try, e.g. 96:1 128:2:4")
    a = ap.parse_args()
    lib = _cabi.load()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6573.8))
    x_rows = 128
    for name in a.shapes:
        n, k = SHAPES_7B[name]
        layers = a.layers if name != "lm_head" else max(2, a.layers // 2)
        ws = [(torch.randn(n, k, device="cuda") * 0.02).to(torch.bfloat16) for _ in range(layers)]
        x = torch.randn(x_rows, k, device="cuda").to(torch.bfloat16)
        c = torch.empty(x_rows, n, dtype=torch.bfloat16, device="cuda")
        bytes_per = n * k * 2 + a.m * k * 2 + a.m * n * 2

        def run_torch(s):
            for w in ws:
                torch.mm(x[:a.m], w.t(), out=c[:a.m])

        ms = time_graph(run_torch, a.reps)
        us = ms * 1e3 / layers
        print(json.dumps({"shape": name, "impl": "torch.mm", "m": a.m, "n": n, "k": k, "us": round(us, 2),
                          "gbs": round(bytes_per / us / 1e3, 1), "frac_hbm": round(bytes_per / us / 1e3 / hbm, 3)}), flush=True)
        for cfg in ["0:0"] + a.configs:
            parts = [int(v) for v in cfg.split(":")]
            tn, sk = parts[0] | ((parts[2] if len(parts) > 2 else 0) << 16) | ((parts[3] if len(parts) > 3 else 0) << 20), parts[1]

            def run_own(s):
                for w in ws:
                    _cabi.check(lib.lade_gemm_bf16(s.cuda_stream, x.data_ptr(), w.data_ptr(), c.data_ptr(), a.m, x_rows, n, k, n,
                                                   tn, sk))
            try:
                ms = time_graph(run_own, a.reps)
            except Exception as e:  # unsupported config
                print(json.dumps({"shape": name, "impl": "lade_gemm_bf16", "cfg": cfg, "error": str(e)[:80]}), flush=True)
                continue
            us = ms * 1e3 / layers
            ref = (x[:a.m].float() @ ws[-1].float().t())
            err = (c[:a.m].float() - ref).abs().max().item()
            print(json.dumps({"shape": name, "impl": "lade_gemm_bf16", "cfg": cfg, "m": a.m, "n": n, "k": k, "us": round(us, 2),
                              "gbs": round(bytes_per / us / 1e3, 1), "frac_hbm": round(bytes_per / us / 1e3 / hbm, 3),
                              "max_abs_err": round(err, 4)}), flush=True)
        del ws


if __name__ == "__main__":
    main()
