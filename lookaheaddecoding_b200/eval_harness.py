"""Throughput harness shaped like the reference's evaluation scripts (``applications/eval_mtbench.py:271-323,384-389``;
``eval_humaneval.py`` / ``eval_cnndm.py`` run the same timing loop over single-turn prompts).

The reference iterates over dataset questions, feeds every turn through ``model.generate`` (lookahead decoding
enabled by ``lade.augment_all(); lade.config_lade(...)``), times each call with the wall clock, accumulates
``overall_time / overall_gen / overall_tp / count_gen`` and ends with ``lade.log_history()`` and
``lade.save_log()``.  Datasets, tokenizers and chat templates are out of scope here (no network, SURVEY 8 "next" row
4 asks for the loop over synthetic prompts): a *question* is a list of token-id turns; the conversation prompt of turn
j is everything said so far (earlier turns and the model's answers) followed by turn j's tokens.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch


@dataclass
class EvalReport:
    overall_time: float = 0.0
    overall_gen: int = 0
    overall_tp: float = 0.0
    count_gen: int = 0
    stats: Dict[int, Dict[int, List[float]]] = field(default_factory=dict)   # question -> turn -> [seconds, tokens]

    @property
    def throughput_mean_of_calls(self) -> float:        # "AVERAGE THROUGHPUT1", eval_mtbench.py:387
        return self.overall_tp / max(self.count_gen, 1)

    @property
    def throughput_overall(self) -> float:              # "AVERAGE THROUGHPUT2"
        return self.overall_gen / max(self.overall_time, 1e-12)

    def summary(self) -> str:
        return (f"AVERAGE THROUGHPUT1 {self.throughput_mean_of_calls} AVERAGE THROUGHPUT2 {self.throughput_overall} "
                f"STAT {[self.overall_tp, self.count_gen, self.overall_gen, self.overall_time]}")


def synthetic_questions(n_questions: int, turns: int, turn_len: int, vocab: int, seed: int = 0,
                        low: int = 3) -> List[List[List[int]]]:
    """`n_questions` conversations of `turns` user turns, `turn_len` random token ids each (MT-bench has 80 x 2)."""
    g = torch.Generator().manual_seed(seed)
    return [[torch.randint(low, vocab, (turn_len,), generator=g).tolist() for _ in range(turns)]
            for _ in range(n_questions)]


@torch.no_grad()
def run_eval(model, questions: Sequence[Sequence[Sequence[int]]], max_new_token: int = 256, temperature: float = 0.0,
             device: Optional[torch.device] = None, max_context: Optional[int] = None, verbose: bool = False,
             sync=None) -> EvalReport:
    """The timing loop of eval_mtbench.py:271-323: per turn one ``model.generate`` (greedy when temperature < 1e-4,
    else sampling with top_k=0, top_p=1.0), wall-clock timed; the answer is appended to the conversation."""
    dev = device or next(model.parameters()).device
    sync = sync or (lambda: torch.cuda.synchronize(dev) if dev.type == "cuda" else None)
    rep = EvalReport()
    for qi, question in enumerate(questions):
        rep.stats[qi] = {}
        torch.manual_seed(0)                                    # `torch.manual_seed(i)` per choice, num_choices = 1
        conversation: List[int] = []
        for ti, turn in enumerate(question):
            conversation = conversation + [int(t) for t in turn]
            if max_context is not None and len(conversation) + max_new_token > max_context:
                conversation = conversation[-(max_context - max_new_token):]      # keep the tail, like a chat window
            ids = torch.tensor([conversation], dtype=torch.long, device=dev)
            do_sample = temperature >= 1e-4
            kw = dict(do_sample=do_sample, max_new_tokens=max_new_token, attention_mask=torch.ones_like(ids))
            if do_sample:
                kw.update(temperature=temperature, top_k=0, top_p=1.0)
            sync()
            t0 = time.time()
            out = model.generate(ids, **kw)
            sync()
            gap = time.time() - t0
            tokens = out.numel() - ids.shape[1]
            rep.overall_time += gap
            rep.overall_gen += tokens
            rep.overall_tp += tokens / gap
            rep.count_gen += 1
            rep.stats[qi][ti] = [gap, tokens]
            if verbose:
                print([f"step {qi} turn {ti} time: ", gap, " generated tokens: ", tokens, " throughput: ", tokens / gap])
            conversation = out[0].tolist()
    return rep
