# round 2 (session 2), 1 GPU: W fragments prepared once per target + one TMA bulk copy per CTA in the v4 kernel, A/B against the previous commit
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2zf_pytest.log 2>&1; tail -3 gpurun_out/r2zf_pytest.log
bash tools/ab_bench.sh DEFAULT prev DEFAULT prev > gpurun_out/r2zf_ab.txt 2>&1; cat gpurun_out/r2zf_ab.txt
timeout 300 python tools/bench_configs.py demc100 > gpurun_out/r2zf_demc.txt 2>&1; grep "^{" gpurun_out/r2zf_demc.txt | cut -c1-300
