# round 2 (session 2), 1 GPU: the full GPU suite once more after the fix of the interop test's generation count
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2zl_pytest.log 2>&1; tail -4 gpurun_out/r2zl_pytest.log
