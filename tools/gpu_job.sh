# round 2, GPU job 3 (1 GPU): v4 kernel (TMA-staged gathers): parity suite, bench A/B against v3, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; tail -15 gpurun_out/r2c_pytest.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2c_bench_v4.json 2> gpurun_out/r2c_bench_v4.err; tail -c 1200 gpurun_out/r2c_bench_v4.json; tail -3 gpurun_out/r2c_bench_v4.err
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary --fused 5 > gpurun_out/r2c_bench_v3.json 2> gpurun_out/r2c_bench_v3.err; tail -c 1200 gpurun_out/r2c_bench_v3.json
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary --history none --burnin-gen 0 > gpurun_out/r2c_bench_v4_plain.json 2>/dev/null; tail -c 900 gpurun_out/r2c_bench_v4_plain.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_gauss_v4 -s 130 -c 1 -o gpurun_out/prof_r2c -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2c_ncu_full.log 2>&1; tail -2 gpurun_out/r2c_ncu_full.log
