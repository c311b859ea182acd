# round 2, GPU job 9 (1 GPU): fly off again; CR fold A/B; checked build over every kernel family
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; tail -6 gpurun_out/r2i_pytest.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2i_bench_fold.json 2> gpurun_out/r2i_bench.err; tail -c 900 gpurun_out/r2i_bench_fold.json; tail -3 gpurun_out/r2i_bench.err
BIPYMC_B200_NO_CR_FOLD=1 timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2i_bench_nofold.json 2>> gpurun_out/r2i_bench.err; tail -c 900 gpurun_out/r2i_bench_nofold.json
timeout 300 python tools/bench_configs.py c3 c4 demc100 > gpurun_out/r2i_secondary.txt 2>&1; cat gpurun_out/r2i_secondary.txt
BIPYMC_B200_LIB=$PWD/build_ab/lib_checks.so timeout 300 python tools/sanitize_case.py > gpurun_out/r2i_checked_build.log 2>&1; tail -6 gpurun_out/r2i_checked_build.log
BIPYMC_B200_FLY=1 timeout 600 python -m pytest tests -m gpu -q -x -k "native or variants or full_size or small_d" > gpurun_out/r2i_pytest_fly.log 2>&1; tail -4 gpurun_out/r2i_pytest_fly.log
