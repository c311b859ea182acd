#!/bin/bash
# A/B kernel builds: run the headline bench once per alternative library (build_ab/lib_*.so);
# the name DEFAULT stands for the in-tree library
for lib in "$@"; do
  for cfg in "" "--history none --burnin-gen 0"; do
    if [ "$lib" = DEFAULT ]; then unset BIPYMC_B200_LIB; else export BIPYMC_B200_LIB=$PWD/build_ab/lib_$lib.so; fi
    python bench.py --steps 50 --warmup 3 --no-cpu --no-e2e $cfg 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib','[$cfg]','%.4g'%j['value'], 'launch_us %.1f'%(1e3*j['roofline']['avg_launch_ms']), 'frac %.3f'%j['roofline']['frac'], 'acc %.5f'%j['acceptance_fraction'])"
  done
done
