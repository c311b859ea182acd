# round 2 (session 2), 1 GPU: re-measure C3 (one run of the artefact job showed 2.8 ms per generation) -- current library against the previous commit's
set -x
mkdir -p gpurun_out
for i in 1 2 3; do
timeout 200 python tools/bench_configs.py c3 > gpurun_out/r2zj_c3_new_$i.txt 2>&1; grep "^{" gpurun_out/r2zj_c3_new_$i.txt | cut -c1-260
done
BIPYMC_B200_LIB=$PWD/build_ab/lib_prev.so timeout 200 python tools/bench_configs.py c3 > gpurun_out/r2zj_c3_prev.txt 2>&1; grep "^{" gpurun_out/r2zj_c3_prev.txt | cut -c1-260
timeout 300 python tools/bench_configs.py c3 c4 demc100 c5shape > gpurun_out/r2zj_all.txt 2>&1; grep "^{" gpurun_out/r2zj_all.txt | cut -c1-260
BIPYMC_B200_LIB=$PWD/build_ab/lib_checks.so timeout 300 python tools/sanitize_case.py > gpurun_out/r2zj_checked_build.log 2>&1; tail -3 gpurun_out/r2zj_checked_build.log
