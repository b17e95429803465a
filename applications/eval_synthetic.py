#!/usr/bin/env python
"""MT-bench-shaped throughput run over synthetic multi-turn prompts (the reference's applications/eval_mtbench.py
timing loop without datasets / tokenizers): random-init Llama of a chosen shape, lookahead decoding through the plugin
surface, per-turn wall-clock timing, lade.log_history() totals at the end.

  USE_LADE=1 python applications/eval_synthetic.py --workload 7b --questions 8 --turns 2 --max-new-token 128
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="7b", choices=["7b", "13b", "tiny"])
    ap.add_argument("--questions", type=int, default=8)
    ap.add_argument("--turns", type=int, default=2)
    ap.add_argument("--turn-len", type=int, default=96)
    ap.add_argument("--max-new-token", type=int, default=128)
    ap.add_argument("--temperature", type=float, default=0.0)
    ap.add_argument("--level", type=int, default=None)
    ap.add_argument("--window", type=int, default=None)
    ap.add_argument("--guess", type=int, default=None)
    ap.add_argument("--use-pp", type=int, default=0, help="accepted for command-line compatibility; ignored")
    ap.add_argument("--save", default=None, help="path for lade.save_log()")
    a = ap.parse_args()

    import torch
    import bench
    import lade
    from lookaheaddecoding_b200.eval_harness import run_eval, synthetic_questions

    shape, W, N, G, _ = bench.WORKLOADS[a.workload]
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    model = bench.build_model(shape, dev)
    os.environ.setdefault("USE_LADE", "1")
    lade.augment_all()
    lade.config_lade(LEVEL=a.level or N, WINDOW_SIZE=a.window or W, GUESS_SET_SIZE=a.guess or G, DEBUG=1, POOL_FROM_PROMPT=True)
    qs = synthetic_questions(a.questions, a.turns, a.turn_len, shape["vocab"])
    run_eval(model, qs[:1], max_new_token=8, temperature=a.temperature, device=dev)          # warm-up (graphs, caches)
    lade.log_history(clear=True)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):        # the per-generate DEBUG summary is kept in the log, not printed
        rep = run_eval(model, qs, max_new_token=a.max_new_token, temperature=a.temperature, device=dev,
                       max_context=shape["max_pos"])
    print(rep.summary())
    lade.log_history()
    if a.save:
        lade.save_log(a.save)
    print(json.dumps({"questions": a.questions, "turns": a.turns, "generate_calls": rep.count_gen,
                      "tokens": rep.overall_gen, "seconds": round(rep.overall_time, 3),
                      "tokens_per_s_overall": round(rep.throughput_overall, 2),
                      "tokens_per_s_mean_of_calls": round(rep.throughput_mean_of_calls, 2)}))


if __name__ == "__main__":
    main()
