# round 2 (session 2), 1 GPU: compile-time DE-MC specialisation of the v4 kernel, A/B against the previous commit's library
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2zc_pytest.log 2>&1; tail -3 gpurun_out/r2zc_pytest.log
for i in 1 2; do
timeout 300 python tools/bench_configs.py demc100 > gpurun_out/r2zc_new_$i.txt 2>&1; grep "^{" gpurun_out/r2zc_new_$i.txt | cut -c1-420
BIPYMC_B200_LIB=$PWD/build_ab/lib_head.so timeout 300 python tools/bench_configs.py demc100 > gpurun_out/r2zc_head_$i.txt 2>&1; grep "^{" gpurun_out/r2zc_head_$i.txt | cut -c1-420
done
