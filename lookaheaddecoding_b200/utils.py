"""Plugin surface of the reference (``lade/utils.py``): config + reversible HF patching.

``config_lade`` keeps the reference's signature verbatim (lade/utils.py:13).  ``augment_generate``
installs the proxies on ``transformers.GenerationMixin`` (transformers 5.x: ``_sample``; 4.36-era
attribute names are patched too when they exist); unlike the reference's destructive
``inject_module`` (lade/utils.py:40-52) nothing in the HF Llama classes is overwritten -- the model
adapter lives in the CUDA engine -- and ``restore_generate`` undoes the patch.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
from transformers import GenerationMixin

from .decoding import CONFIG_MAP, FUNC_MAP, greedy_search_proxy, sample_entry_proxy, sample_proxy


# Knobs that are plain CONFIG_MAP entries.  SPLIT_FLAG is a dead knob in the reference too (lade/utils.py:24-25);
# USE_FLASH is accepted and ignored (the fused attention kernel is always on).
_PLAIN_KNOBS = ("WINDOW_SIZE", "LEVEL", "GUESS_SET_SIZE", "ALWAYS_FWD_ONE", "DEBUG", "SPLIT_FLAG", "POOL_FROM_PROMPT",
                "USE_FLASH")


def _join_lookahead_workers(n_workers: int, backend: str) -> None:
    """Lookahead parallelism set-up (lade/utils.py:28-33): one process per GPU, rank = LOCAL_RANK."""
    local_rank = int(os.environ["LOCAL_RANK"])
    CONFIG_MAP.update(DIST_WORKERS=n_workers, LOCAL_RANK=local_rank)
    if not dist.is_initialized():
        dist.init_process_group(backend, rank=local_rank)
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    assert dist.get_world_size() == n_workers, "DIST_WORKERS config should be equal to work size"


def config_lade(WINDOW_SIZE=None, LEVEL=None, DEBUG=None, GUESS_SET_SIZE=None, ALWAYS_FWD_ONE=None, SPLIT_FLAG=None,
                DIST_WORKERS=None, POOL_FROM_PROMPT=None, backend='nccl', USE_FLASH=None):
    """Same signature and semantics as lade/utils.py:13: every argument that is not None overwrites its CONFIG_MAP
    entry, DIST_WORKERS > 1 joins the process group, and the step log is reset."""
    given = locals()
    CONFIG_MAP.update({knob: given[knob] for knob in _PLAIN_KNOBS if given[knob] is not None})
    if DIST_WORKERS is not None and DIST_WORKERS > 1:
        _join_lookahead_workers(DIST_WORKERS, backend)
    CONFIG_MAP["log"] = []


def augment_llama():
    """The reference injects its modeling_llama methods into HF's classes (utils.py:55-56).  Here the
    model adapter is the CUDA engine, which reads the HF module's weights; nothing to inject.  Kept
    for API parity; loads the native library so a missing build fails here, loudly."""
    from . import _cabi
    _cabi.load()


def augment_generate():
    if "_sample" not in FUNC_MAP and hasattr(GenerationMixin, "_sample"):
        FUNC_MAP["_sample"] = GenerationMixin._sample
        GenerationMixin._sample = sample_entry_proxy
    # transformers 4.36-era entry points (lade/utils.py:62-66), when present
    if hasattr(GenerationMixin, "greedy_search") and "greedy_search" not in FUNC_MAP:
        FUNC_MAP["greedy_search"] = GenerationMixin.greedy_search
        FUNC_MAP["sample"] = GenerationMixin.sample
        GenerationMixin.greedy_search = greedy_search_proxy
        GenerationMixin.sample = sample_proxy


def restore_generate():
    if "_sample" in FUNC_MAP:
        GenerationMixin._sample = FUNC_MAP.pop("_sample")
    if "greedy_search" in FUNC_MAP:
        GenerationMixin.greedy_search = FUNC_MAP.pop("greedy_search")
        GenerationMixin.sample = FUNC_MAP.pop("sample")


def augment_all():
    augment_llama()
    augment_generate()


def log_history(clear=False):
    """Totals over the (generated tokens, steps) pairs logged by DEBUG runs; same report line as lade/utils.py:74-83."""
    records = CONFIG_MAP.get("log", [])
    gen, step = sum(r[0] for r in records), sum(r[1] for r in records)
    if clear:
        CONFIG_MAP["log"] = []
    ratio = (gen / step) if step > 0 else 0
    print("LADE LOG - OVERALL GEN: ", gen, " STEPS: ", step, " AVG COMPRESS RATIO: ", ratio)


def save_log(log_dir):
    """torch.save of the step log, if any (lade/utils.py:85-87)."""
    records = CONFIG_MAP.get("log")
    if records is not None:
        torch.save(records, log_dir)
