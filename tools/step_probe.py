#!/usr/bin/env python
"""How much of a decode step is GPU work?  Replays the captured steady-step CUDA graph back to back (no host in the
loop) and compares with the per-step time of a normal generate() (one host sync per step)."""
import json
import os
import random
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from lookaheaddecoding_b200 import LookaheadEngine  # noqa: E402


@torch.no_grad()
def main():
    shape = bench.WORKLOADS["7b"][0]
    dev = torch.device("cuda", 0)
    model = bench.build_model(shape, dev)
    W, N, G, P, new = 15, 5, 15, 1024, 256
    eng = LookaheadEngine(model, W, N, G, pool_from_prompt=True, max_total_len=P + new + 8)
    torch.manual_seed(1)
    prompt = torch.randint(3, shape["vocab"], (P,)).tolist()
    for _ in range(2):
        eng.generate(prompt, new, rng=random.Random(0))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.generate(prompt, new, rng=random.Random(0))
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    steps = eng.last_steps
    gen_ms = e0.elapsed_time(e1)
    # graph-only: restart a generate, run the eager prefill + fill steps, then replay the steady graph with no host sync
    eng.begin(prompt, P + new, (), eng.draw_window(prompt, random.Random(0), None))
    for s in range(N - 2):
        eng.run_forward_step(s, P)
        eng._read_result()
    g = eng._steady_graph(True)
    reps = 100
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"generate_ms_per_step": round(gen_ms / steps, 4), "generate_wall_ms_per_step": round(wall * 1e3 / steps, 4),
                      "steps": steps, "graph_only_ms_per_step": round(graph_ms, 4),
                      "host_gap_us_per_step": round((gen_ms / steps - graph_ms) * 1e3, 1)}))


if __name__ == "__main__":
    main()
