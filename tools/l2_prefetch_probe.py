#!/usr/bin/env python
"""Does a cuBLAS projection GEMM run faster when (part of) its weights were prefetched into L2 while another kernel ran?

For each 7B projection shape: time  [filler kernel ; GEMM]  with and without a concurrent lade_l2_prefetch of the GEMM's
weights on a side stream during the filler (the filler stands for attention/norm phases: it keeps the SMs busy but HBM
idle).  Weights rotate over enough distinct sets that nothing is L2-resident by accident."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lookaheaddecoding_b200 import _cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=120)
    ap.add_argument("--sets", type=int, default=8)
    ap.add_argument("--filler-us", type=float, nargs="+", default=[6.0, 16.0])
    ap.add_argument("--ctas", type=int, nargs="+", default=[16, 64])
    ap.add_argument("--chunk", type=int, nargs="+", default=[16384, 65536])
    ap.add_argument("--frac", type=float, nargs="+", default=[1.0])
    a = ap.parse_args()
    lib = _cabi.load()
    dev = "cuda"
    shapes = {"qkv": (12288, 4096), "o": (4096, 4096), "gate_up": (22016, 4096), "down": (4096, 11008)}
    main_s = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    # filler: a compute-only kernel of tunable length (keeps SMs busy, touches no DRAM): small matmul chain in L2/smem
    fa = torch.randn(2048, 2048, device=dev, dtype=torch.bfloat16)
    fb = torch.randn(2048, 2048, device=dev, dtype=torch.bfloat16)
    fc = torch.empty(2048, 2048, device=dev, dtype=torch.bfloat16)

    def time_graph(fn, reps=5):
        g = torch.cuda.CUDAGraph()
        cap = torch.cuda.Stream()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            fn()                      # warm-up outside capture
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=cap):
                fn()
        torch.cuda.current_stream().wait_stream(cap)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    # calibrate the filler: n matmuls of 2048^3 (each ~10 us at 1.6 PF)
    def filler(n):
        for _ in range(n):
            torch.mm(fa, fb, out=fc)
    t1 = time_graph(lambda: filler(8)) / 8
    print(json.dumps({"filler_unit_us": round(t1, 2)}))
    for name, (n, k) in shapes.items():
        ws = [torch.randn(n, k, device=dev, dtype=torch.bfloat16) for _ in range(a.sets)]
        x = torch.randn(a.m, k, device=dev, dtype=torch.bfloat16)
        y = torch.empty(a.m, n, device=dev, dtype=torch.bfloat16)
        wbytes = n * k * 2

        def gemm_only():
            for w in ws:
                torch.mm(x, w.t(), out=y)
        t_gemm = time_graph(gemm_only) / a.sets
        for fus in a.filler_us:
            nf = max(1, round(fus / t1))

            def base():
                for w in ws:
                    filler(nf)
                    torch.mm(x, w.t(), out=y)
            t_base = time_graph(base) / a.sets
            for ctas in a.ctas:
                for chunk in a.chunk:
                    for frac in a.frac:
                        nb = int(wbytes * frac) & ~15

                        def with_pf():
                            cur = torch.cuda.current_stream()
                            for w in ws:
                                side.wait_stream(cur)                 # fork: prefetch runs beside the filler
                                _cabi.check(lib.lade_l2_prefetch(side.cuda_stream, w.data_ptr(), nb, ctas, chunk))
                                filler(nf)
                                cur.wait_stream(side)                 # join before the GEMM (the prefetch KERNEL is short)
                                torch.mm(x, w.t(), out=y)
                        t_pf = time_graph(with_pf) / a.sets
                        print(json.dumps({"shape": name, "n": n, "k": k, "MB": round(wbytes / 1e6, 1), "gemm_alone_us": round(t_gemm, 2),
                                          "filler_us": round(nf * t1, 1), "filler+gemm_us": round(t_base, 2),
                                          "with_prefetch_us": round(t_pf, 2), "saved_us": round(t_base - t_pf, 2),
                                          "ctas": ctas, "chunk": chunk, "frac": frac}))
        del ws


if __name__ == "__main__":
    main()
