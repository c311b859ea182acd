# round 2, GPU job 4 (1 GPU): v4.1 (named-barrier hand-over, no divergence in the producers): parity + bench + ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; tail -8 gpurun_out/r2d_pytest.log
timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2d_bench_v4.json 2> gpurun_out/r2d_bench_v4.err; tail -c 1000 gpurun_out/r2d_bench_v4.json; tail -3 gpurun_out/r2d_bench_v4.err
for lib in build_ab/lib_*.so; do
  [ -f "$lib" ] || continue
  BIPYMC_B200_LIB=$PWD/$lib timeout 300 python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2d_bench_$(basename $lib .so).json 2>/dev/null; echo $lib; tail -c 700 gpurun_out/r2d_bench_$(basename $lib .so).json
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_gauss_v4 -s 130 -c 1 -o gpurun_out/prof_r2d -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2d_ncu_full.log 2>&1; tail -2 gpurun_out/r2d_ncu_full.log
BIPYMC_B200_LIB=$PWD/build_ab/lib_zen1.so timeout 600 python -m pytest tests -m gpu -q -x -k "native or variants or full_size or smoke" > gpurun_out/r2d_pytest_zen1.log 2>&1; tail -5 gpurun_out/r2d_pytest_zen1.log
