# round 2, GPU job 10 (1 GPU): v4 without fold / fly hooks (0 spills): parity, bench, secondary, checked build, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; tail -6 gpurun_out/r2j_pytest.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 900 gpurun_out/r2j_bench.json; tail -3 gpurun_out/r2j_bench.err
timeout 300 python tools/bench_configs.py c3 c4 demc100 > gpurun_out/r2j_secondary.txt 2>&1; cat gpurun_out/r2j_secondary.txt
BIPYMC_B200_LIB=$PWD/build_ab/lib_checks.so timeout 300 python tools/sanitize_case.py > gpurun_out/r2j_checked_build.log 2>&1; tail -6 gpurun_out/r2j_checked_build.log
