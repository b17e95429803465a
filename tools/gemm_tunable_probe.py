#!/usr/bin/env python
"""Library GEMM selection: torch.mm (cuBLAS heuristic) vs PyTorch TunableOp (benchmarks the cuBLAS / cuBLASLt algorithm
list once per shape and keeps the fastest) on the decode step's projection shapes, m = 120, graph-replayed over 8
distinct weight sets (> L2)."""
import json
import os
import sys

import torch


def time_graph(fn, reps=5):
    g = torch.cuda.CUDAGraph()
    cap = torch.cuda.Stream()
    cap.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cap):
        fn()
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


def main():
    dev = "cuda"
    shapes = {"qkv": (12288, 4096, 120), "o": (4096, 4096, 120), "gate_up": (22016, 4096, 120), "down": (4096, 11008, 120),
              "lm_head": (32000, 4096, 76), "qkv13": (15360, 5120, 240), "gu13": (27648, 5120, 240)}
    sets = 8
    res = {}
    for mode in ("default", "tunable"):
        if mode == "tunable":
            import torch.cuda.tunable as tn
            tn.enable(True)
            tn.tuning_enable(True)
            tn.set_max_tuning_duration(30)
            tn.set_max_tuning_iterations(100)
            try:
                tn.set_filename(os.path.join("gpurun_out", "tunableop_results.csv"))
            except Exception:
                pass
        for name, (n, k, m) in shapes.items():
            ws = [torch.randn(n, k, device=dev, dtype=torch.bfloat16) for _ in range(sets)]
            x = torch.randn(m, k, device=dev, dtype=torch.bfloat16)
            y = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
            torch.mm(x, ws[0].t(), out=y)           # tuning happens here (eager, outside capture)
            torch.cuda.synchronize()

            def run():
                for w in ws:
                    torch.mm(x, w.t(), out=y)
            us = time_graph(run) / sets
            res.setdefault(name, {})[mode] = round(us, 2)
            res[name]["GBps_" + mode] = round(n * k * 2 / us / 1e3, 1)
            del ws
    for name, r in res.items():
        r["gain_pct"] = round(100 * (r["default"] - r["tunable"]) / r["default"], 1)
        print(json.dumps({"shape": name, **r}))


if __name__ == "__main__":
    main()
