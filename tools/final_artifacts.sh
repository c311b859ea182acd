# Recipe of the round-2 final artefacts under profiles/r2/r2zh_* (one gpurun call on one B200)
set -x
T=r2zh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 2800 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; tail -c 500 gpurun_out/${T}_bench_ref.json
timeout 400 python tools/bench_configs.py c3 c4 demc100 c5shape > gpurun_out/${T}_secondary.txt 2>&1; cat gpurun_out/${T}_secondary.txt
BIPYMC_B200_LIB=$PWD/build_ab/lib_checks.so timeout 300 python tools/sanitize_case.py > gpurun_out/${T}_checked_build.log 2>&1; tail -5 gpurun_out/${T}_checked_build.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/${T}_plain.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/${T}_ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_gauss_v4 -s 130 -c 1 -o gpurun_out/prof_${T} -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-stationary > gpurun_out/${T}_ncu_full.log 2>&1; tail -2 gpurun_out/${T}_ncu_full.log
