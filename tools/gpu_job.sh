# round 2 (session 2), 1 GPU: block-wise 16-byte propose / accept kernels (large d), A/B of the register cap and against the previous commit
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2zb_pytest.log 2>&1; tail -4 gpurun_out/r2zb_pytest.log
timeout 300 python tools/bench_configs.py c4 c3 demc100 c5shape > gpurun_out/r2zb_new.txt 2>&1; grep "^{" gpurun_out/r2zb_new.txt | cut -c1-250
BIPYMC_B200_LIB=$PWD/build_ab/lib_p174.so timeout 300 python tools/bench_configs.py c5shape > gpurun_out/r2zb_p174.txt 2>&1; grep "^{" gpurun_out/r2zb_p174.txt | cut -c1-250
BIPYMC_B200_LIB=$PWD/build_ab/lib_head.so timeout 300 python tools/bench_configs.py c3 c5shape > gpurun_out/r2zb_head.txt 2>&1; grep "^{" gpurun_out/r2zb_head.txt | cut -c1-250
