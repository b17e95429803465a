"""``import lade`` compatibility alias: the reference's package name, served by lookaheaddecoding_b200."""
from lookaheaddecoding_b200 import *  # noqa: F401,F403
from lookaheaddecoding_b200 import decoding, utils, lade_distributed  # noqa: F401
