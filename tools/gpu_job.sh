# round 2, GPU job 6 (1 GPU): persistent small-d kernel, host-peer write-back default, secondary configs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; tail -8 gpurun_out/r2f_pytest.log
timeout 300 python tools/bench_configs.py c3 c4 demc100 > gpurun_out/r2f_secondary.txt 2>&1; cat gpurun_out/r2f_secondary.txt
timeout 300 python tools/bench_configs.py c4multi > gpurun_out/r2f_c4_n1.txt 2>&1; cat gpurun_out/r2f_c4_n1.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-stationary > gpurun_out/r2f_bench_hostpeer.json 2> gpurun_out/r2f_bench.err; tail -c 900 gpurun_out/r2f_bench_hostpeer.json
BIPYMC_B200_HOST_PEER=0 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-stationary > gpurun_out/r2f_bench_nohostpeer.json 2>> gpurun_out/r2f_bench.err; tail -c 900 gpurun_out/r2f_bench_nohostpeer.json
