"""TEST INFRASTRUCTURE ONLY -- the loader of the unmodified reference lives in ``baseline/ref_loader.py``; this
module keeps the old import path of the golden-vector generators (``tests/golden/gen_golden*.py``) working."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline.ref_loader import *          # noqa: F401,F403,E402
from baseline.ref_loader import (REFERENCE_ROOT, build_reference_model, load_reference, make_llama_config,  # noqa: F401,E402
                                 reference_available, run_reference_greedy)
