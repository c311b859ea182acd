# round 2, first GPU pass: parity suite, bench, fair gather probe, host-peer e2e test, ncu of the lazy-protocol kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
timeout 120 tools/gather_probe2 100000 ldg > gpurun_out/r2_gather_probe2.txt 2>&1
timeout 120 tools/gather_probe2 100000 bulk >> gpurun_out/r2_gather_probe2.txt 2>&1
timeout 120 tools/gather_probe2 100000 g4 >> gpurun_out/r2_gather_probe2.txt 2>&1
timeout 120 tools/gather_probe2 400000 bulk >> gpurun_out/r2_gather_probe2.txt 2>&1
cat gpurun_out/r2_gather_probe2.txt
BIPYMC_B200_HOST_PEER=1 timeout 300 python -m pytest tests -m gpu -q -k "host_buffer" > gpurun_out/r2a_hostpeer.log 2>&1; tail -5 gpurun_out/r2a_hostpeer.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_gauss_v3 -s 130 -c 1 -o gpurun_out/prof_r2a -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_ncu_full.log 2>&1; tail -2 gpurun_out/r2a_ncu_full.log
