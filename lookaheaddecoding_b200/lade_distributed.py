"""Device / lookahead-parallelism queries of the plugin surface (same names and meaning as
lade/lade_distributed.py:5-12): which CUDA device this rank decodes on, and whether more than one
lookahead worker was configured with ``config_lade(DIST_WORKERS=...)``."""
from .decoding import CONFIG_MAP


def get_device() -> int:
    """Local rank recorded by config_lade (one process per GPU); 0 for a single-process run."""
    return int(CONFIG_MAP.get("LOCAL_RANK", 0))


def distributed() -> bool:
    """True once config_lade joined a process group of more than one lookahead worker."""
    return int(CONFIG_MAP.get("DIST_WORKERS", 1)) > 1
