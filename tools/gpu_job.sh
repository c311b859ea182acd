# round 2, GPU job 8 (1 GPU): fly mode (no split kernel), CR reduction folded into the v4 tail, fly small-d kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; tail -8 gpurun_out/r2h_pytest.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 1300 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err
timeout 300 python tools/bench_configs.py c3 c4 demc100 > gpurun_out/r2h_secondary.txt 2>&1; cat gpurun_out/r2h_secondary.txt
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2h_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2h_ncu_l.log 2>&1
