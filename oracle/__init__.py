"""TEST INFRASTRUCTURE ONLY.  CPU oracle for the lookahead/verification decoding step.

``oracle.lookahead``  -- integer state machine (window, n-gram pool, layout, mask predicate, accept)
``oracle.llama_ref``  -- torch restatement of the floating-point step (reference eager numerics)
``oracle.ref_shim``   -- loader of the unmodified reference from /root/reference (fixture generation)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
