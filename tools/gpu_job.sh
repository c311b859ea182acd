# round 2, GPU job 5 (1 GPU): default = v4 + single-Philox draws; full parity suite, full bench line, secondary
# configs, C5 plumbing at small scale, sanitizer passes, ncu artefacts
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -8 gpurun_out/r2e_pytest.log
timeout 600 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 2600 gpurun_out/r2e_bench.json; tail -3 gpurun_out/r2e_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_bench_ref.json 2> gpurun_out/r2e_bench_ref.err; tail -c 600 gpurun_out/r2e_bench_ref.json
timeout 300 python tools/bench_configs.py c3 c4 demc100 > gpurun_out/r2e_secondary.txt 2>&1; cat gpurun_out/r2e_secondary.txt
C5_PER_GPU=40000 C5_GENS=5 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29561 tools/bench_configs.py c5full > gpurun_out/r2e_c5small.txt 2>&1; tail -3 gpurun_out/r2e_c5small.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r2e_sanitizer_memcheck.log 2>&1; echo memcheck rc=$?; tail -4 gpurun_out/r2e_sanitizer_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r2e_sanitizer_racecheck.log 2>&1; echo racecheck rc=$?; tail -4 gpurun_out/r2e_sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_case.py > gpurun_out/r2e_sanitizer_synccheck.log 2>&1; echo synccheck rc=$?; tail -4 gpurun_out/r2e_sanitizer_synccheck.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2e_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2e_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_gauss_v4 -s 130 -c 1 -o gpurun_out/prof_r2e -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/r2e_ncu_full.log 2>&1; tail -2 gpurun_out/r2e_ncu_full.log
