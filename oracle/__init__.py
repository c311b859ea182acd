"""CPU oracle for the DE-MC / DREAM hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import anything under oracle/.  The product package (bipymc_b200/) must not,
and fails loudly when its CUDA library is missing instead of falling back to this.
Parity pin: tests/test_oracle_golden.py checks oracle/demc_dream.py bit-for-bit against
tests/golden/ref_*.npz, which oracle/make_golden.py produced by running the unmodified
reference (/root/reference) in the build container.
"""
