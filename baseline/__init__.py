"""Reference-side tooling (loader of the unmodified reference, GPU-vs-GPU parity harness).  Not product code."""
