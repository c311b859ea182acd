"""Single-rank stand-in for mpi4py, used ONLY to import the unmodified reference
under /root/reference when generating golden vectors (oracle/make_golden.py).
Test infrastructure; never imported by the product package."""
from . import MPI  # noqa: F401
