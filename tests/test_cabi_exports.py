"""CPU-only: the C-ABI library builds/loads and exports every symbol include/lade_sm100.h declares."""
import ctypes as C
import os
import re

from lookaheaddecoding_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported_and_bound():
    with open(os.path.join(ROOT, "include", "lade_sm100.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(lade_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _cabi.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_cabi.EXPORTED_SYMBOLS), declared ^ set(_cabi.EXPORTED_SYMBOLS)


def test_host_only_entry_points():
    lib = _cabi.load()
    assert lib.lade_version() >= 100
    assert lib.lade_strerror(0) == b"ok"
    c = _cabi.LadeConfig()
    c.window_size, c.level, c.guess_set_size = 15, 5, 15
    assert lib.lade_step_rows_bound(C.byref(c), 64, 0) == 64 + 17      # P + W+N-3   (P+17)
    assert lib.lade_step_rows_bound(C.byref(c), 64, 1) == 34
    assert lib.lade_step_rows_bound(C.byref(c), 64, 2) == 48
    assert lib.lade_step_rows_bound(C.byref(c), 64, 3) == 120
    c.window_size, c.level, c.guess_set_size = 20, 7, 20
    assert lib.lade_step_rows_bound(C.byref(c), 1, 9) == 240
    assert lib.lade_attn_scratch_bytes(120, 32, 128, 5) == 16384 * 4 + 5 * 32 * 128 * 130 * 4
    # bad arguments are rejected with codes, never crashes (no GPU needed: checks precede any CUDA call)
    assert lib.lade_ctx_create(None, None) == -1
    bad = _cabi.LadeConfig()
    bad.window_size, bad.level, bad.guess_set_size, bad.vocab_size, bad.max_total_len = 5, 2, 3, 100, 10
    out = C.c_void_p()
    assert lib.lade_ctx_create(C.byref(bad), C.byref(out)) == -1          # LEVEL < 3
    bad.level, bad.guess_set_size = 3, -1
    assert lib.lade_ctx_create(C.byref(bad), C.byref(out)) == -4          # unbounded pool unsupported
    assert lib.lade_attn_fwd(None, None, None, None, None, None, 0, None, None, 1, 1, 1, 128, 1, 1, 1, 0) == -1


def test_plugin_surface_signatures():
    import inspect
    import lade
    sig = inspect.signature(lade.config_lade)
    assert list(sig.parameters) == ["WINDOW_SIZE", "LEVEL", "DEBUG", "GUESS_SET_SIZE", "ALWAYS_FWD_ONE", "SPLIT_FLAG",
                                    "DIST_WORKERS", "POOL_FROM_PROMPT", "backend", "USE_FLASH"]   # lade/utils.py:13
    assert sig.parameters["backend"].default == "nccl"
    for name in ("augment_all", "augment_llama", "augment_generate", "log_history", "save_log", "get_device", "distributed"):
        assert callable(getattr(lade, name))
    lade.config_lade(LEVEL=5, WINDOW_SIZE=15, GUESS_SET_SIZE=15, DEBUG=0)
    from lookaheaddecoding_b200.decoding import CONFIG_MAP
    assert CONFIG_MAP["LEVEL"] == 5 and CONFIG_MAP["log"] == []
    assert lade.get_device() == 0 and not lade.distributed()
