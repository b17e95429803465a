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


def config_lade(WINDOW_SIZE=None, LEVEL=None, DEBUG=None, GUESS_SET_SIZE=None, ALWAYS_FWD_ONE=None, SPLIT_FLAG=None,
                DIST_WORKERS=None, POOL_FROM_PROMPT=None, backend='nccl', USE_FLASH=None):
    if WINDOW_SIZE is not None:
        CONFIG_MAP["WINDOW_SIZE"] = WINDOW_SIZE
    if LEVEL is not None:
        CONFIG_MAP["LEVEL"] = LEVEL
    if GUESS_SET_SIZE is not None:
        CONFIG_MAP["GUESS_SET_SIZE"] = GUESS_SET_SIZE
    if ALWAYS_FWD_ONE is not None:
        CONFIG_MAP["ALWAYS_FWD_ONE"] = ALWAYS_FWD_ONE
    if DEBUG is not None:
        CONFIG_MAP["DEBUG"] = DEBUG
    if SPLIT_FLAG is not None:
        CONFIG_MAP["SPLIT_FLAG"] = SPLIT_FLAG          # dead knob in the reference too (utils.py:24-25)
    if POOL_FROM_PROMPT is not None:
        CONFIG_MAP["POOL_FROM_PROMPT"] = POOL_FROM_PROMPT
    if DIST_WORKERS is not None and DIST_WORKERS > 1:
        CONFIG_MAP["DIST_WORKERS"] = DIST_WORKERS
        CONFIG_MAP["LOCAL_RANK"] = int(os.environ["LOCAL_RANK"])
        if not dist.is_initialized():
            dist.init_process_group(backend, rank=CONFIG_MAP["LOCAL_RANK"])
        if torch.cuda.is_available():
            torch.cuda.set_device(CONFIG_MAP["LOCAL_RANK"])
        assert dist.get_world_size() == DIST_WORKERS, "DIST_WORKERS config should be equal to work size"
    if USE_FLASH is not None:
        CONFIG_MAP["USE_FLASH"] = USE_FLASH            # accepted, ignored: the fused kernel is always on
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
    gen = 0
    step = 0
    if "log" in CONFIG_MAP:
        for log in CONFIG_MAP["log"]:
            gen += log[0]
            step += log[1]
    if clear:
        CONFIG_MAP["log"] = []
    print("LADE LOG - OVERALL GEN: ", gen, " STEPS: ", step, " AVG COMPRESS RATIO: ", (gen / step) if step > 0 else 0)


def save_log(log_dir):
    if "log" in CONFIG_MAP:
        torch.save(CONFIG_MAP["log"], log_dir)
