#!/usr/bin/env python
"""Golden traces of the reference's lookahead-parallel (LP, DIST_WORKERS>1) greedy loop.

Spawns D CPU processes (gloo), each running the UNMODIFIED reference loop
(lade/decoding.py:697, LP branches :905-906,:956-984,:1023-1024,:1043-1058,:1088-1107,:1148-1153) on the same
seeded tiny model, and records per rank and per step what went into / came out of jforward_multilevel.
Output: tests/golden/lp_traces.json.gz.   Re-run: python tests/golden/gen_golden_lp.py
"""
from __future__ import annotations

import gzip
import json
import os
import sys
import tempfile
import warnings

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
warnings.filterwarnings("ignore")

TINY = dict(hidden=256, layers=2, heads=2, inter=688, vocab=32000, max_pos=2048)
LP_CASES = [
    # name, dtype, W, N, G, pool, D, prompt_len, max_new, weight seed, prompt seed
    ("lp2_fp32_w15n5g15_pool", "float32", 15, 5, 15, True, 2, 48, 96, 0, 3),
    ("lp3_fp32_w7n4g6", "float32", 7, 4, 6, True, 3, 32, 64, 0, 9),
    ("lp4_fp32_w5n3g3_pool", "float32", 5, 3, 3, True, 4, 24, 48, 0, 1),
    ("lp2_bf16_w15n5g15_pool", "bfloat16", 15, 5, 15, True, 2, 64, 96, 1, 4),
    ("lp8_fp32_w15n5g15_pool", "float32", 15, 5, 15, True, 8, 40, 64, 0, 6),
]


def _worker(rank, D, case, init_file, out_dir):
    from oracle import ref_shim as R
    sys.path.insert(0, HERE)
    from gen_golden import trace_reference_greedy

    (name, dt, W, N, G, pool, _, P, max_new, wseed, pseed) = case
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=D)
    torch.set_num_threads(2)
    cfg = R.make_llama_config(**TINY)
    model = R.build_reference_model(cfg, seed=wseed, dtype=getattr(torch, dt))
    torch.manual_seed(pseed)
    prompt = torch.randint(3, TINY["vocab"], (1, P))
    lade_cfg = dict(WINDOW_SIZE=W, LEVEL=N, GUESS_SET_SIZE=G, DEBUG=1, POOL_FROM_PROMPT=int(pool),
                    DIST_WORKERS=D, LOCAL_RANK=rank)
    # every rank seeds python's RNG differently on purpose: rank 0's window is broadcast (decoding.py:906)
    ids, steps, pool_d, log, _ = trace_reference_greedy(model, prompt, max_new, lade_cfg, py_seed=pseed + 100 * rank)
    for s in steps:
        s.pop("mask_rows", None) if False else None
    with open(os.path.join(out_dir, f"{name}.{rank}.json"), "w") as f:
        json.dump(dict(output_ids=ids, steps=steps, final_pool=pool_d, n_steps=log[1] if rank == 0 else len(steps)), f)
    dist.barrier()
    dist.destroy_process_group()


def main():
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for case in LP_CASES:
            name, D = case[0], case[6]
            init_file = os.path.join(td, f"init_{name}")
            mp.spawn(_worker, args=(D, case, init_file, td), nprocs=D, join=True)
            ranks = []
            for r in range(D):
                with open(os.path.join(td, f"{name}.{r}.json")) as f:
                    ranks.append(json.load(f))
            (_, dt, W, N, G, pool, _, P, max_new, wseed, pseed) = case
            torch.manual_seed(pseed)
            prompt = torch.randint(3, TINY["vocab"], (1, P))[0].tolist()
            assert all(rk["output_ids"] == ranks[0]["output_ids"] for rk in ranks)
            out[name] = dict(model=TINY, dtype=dt, W=W, N=N, G=G, pool_from_prompt=bool(pool), D=D, weight_seed=wseed,
                             prompt=prompt, max_new=max_new, py_seed=pseed, output_ids=ranks[0]["output_ids"],
                             n_steps=len(ranks[0]["steps"]), ranks=ranks)
            print(f"{name}: D={D} generated {len(ranks[0]['output_ids']) - P} tokens in {len(ranks[0]['steps'])} steps")
    with gzip.open(os.path.join(HERE, "lp_traces.json.gz"), "wt") as f:
        json.dump(out, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
