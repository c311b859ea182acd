# round 2, 1-GPU A/B: line-fit likelihood summation form (four interleaved partials vs the sequential sum unrolled by five)
set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_configs.py c4 c3 > gpurun_out/r2s_c4_partials.txt 2>&1; cat gpurun_out/r2s_c4_partials.txt
BIPYMC_B200_LIB=$PWD/build_ab/lib_linefitseq.so timeout 300 python tools/bench_configs.py c4 > gpurun_out/r2s_c4_seq.txt 2>&1; cat gpurun_out/r2s_c4_seq.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "linefit or small_d or replay" > gpurun_out/r2s_pytest.log 2>&1; tail -3 gpurun_out/r2s_pytest.log
