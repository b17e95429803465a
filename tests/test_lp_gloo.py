"""Lookahead parallelism (DIST_WORKERS>1) on CPU: the oracle's LP restatement run as real world_size-D gloo
processes must reproduce, rank by rank and step by step, the traces of the unmodified reference run the same
way (tests/golden/lp_traces.json.gz, produced by tests/golden/gen_golden_lp.py)."""
import gzip
import json
import os
import random
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load_lp_cases():
    with gzip.open(os.path.join(GOLD, "lp_traces.json.gz"), "rt") as f:
        return json.load(f)


LP_CASES = load_lp_cases()


def _worker(rank, D, name, init_file, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import lookahead as LA
    from oracle import llama_ref as LR

    c = load_lp_cases()[name]
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=D)
    torch.set_num_threads(2)
    w = LR.init_weights(c["model"], seed=c["weight_seed"], dtype=getattr(torch, c["dtype"]))
    om = LR.OracleLlama(c["model"], w)
    trace, pool = [], {}
    ids, steps = LA.greedy_lookahead(c["prompt"], c["max_new"], c["W"], c["N"], c["G"], om.step_fn, om.compact_fn,
                                     pool_from_prompt=c["pool_from_prompt"],
                                     rng=random.Random(c["py_seed"] + 100 * rank), trace=trace, token_map_out=pool,
                                     comm=LA.TorchDistComm())
    rec = dict(ids=ids, steps=steps,
               trace=[dict(ids=t.ids, pos=t.pos, guess=t.guess_tokens, first=t.first_guess, inp=t.inp_tokens,
                           gres=t.guess_results, kv_len=t.kv_len) for t in trace],
               pool={str(k): [list(t) for t in v] for k, v in pool.items()})
    with open(os.path.join(out_dir, f"{rank}.json"), "w") as f:
        json.dump(rec, f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", sorted(LP_CASES))
def test_oracle_lp_matches_reference_lp(name):
    c = LP_CASES[name]
    D = c["D"]
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(D, name, os.path.join(td, "init"), td), nprocs=D, join=True)
        for r in range(D):
            with open(os.path.join(td, f"{r}.json")) as f:
                got = json.load(f)
            ref = c["ranks"][r]
            assert got["ids"] == c["output_ids"]
            assert got["steps"] == len(ref["steps"])
            for i, (t, g) in enumerate(zip(got["trace"], ref["steps"])):
                n_in = len(g["input_ids"])
                flat = list(g["input_ids"])
                for lvl in g["past_tokens"][: g["fill_level"] + 1]:
                    flat = flat + lvl
                flat = flat + (g["guess_tokens"] or [])
                assert t["ids"] == flat, f"rank {r} step {i} rows"
                assert t["pos"][:n_in] == g["position_ids"], f"rank {r} step {i} positions"
                assert t["guess"] == g["guess_tokens"]
                assert t["first"] == g["first_guess"] and t["inp"] == g["inp_tokens"] and t["gres"] == g["guess_results"]
                assert t["kv_len"] + n_in == g["kvcache_len"]
            assert got["pool"] == ref["final_pool"]
