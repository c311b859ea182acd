set -x
python bench.py > gpurun_out/r1b_bench.json 2> gpurun_out/r1b_bench.err; tail -c 600 gpurun_out/r1b_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1b_bench_ref.json 2> gpurun_out/r1b_bench_ref.err; tail -c 400 gpurun_out/r1b_bench_ref.json
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r1b_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/r1b_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_gauss_v3 -s 130 -c 1 -o gpurun_out/prof_r1b_final -f python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r1b_ncu_full.log 2>&1; tail -2 gpurun_out/r1b_ncu_full.log
