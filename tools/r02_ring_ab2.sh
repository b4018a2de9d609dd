#!/bin/bash
# scratch ring x eager job taking by the worker warps: throughput and DRAM traffic on C5-2048.  Needs a library built with
# ABCOCT_BUILD_RING=1; ABCOCT_HINTS bit 8 (workers take any ready job, no backlog gate) existed only for this A/B and was removed
# again after it showed no effect (profiles/r02_scratch_ring_experiments.txt, part 3).
set -u
mkdir -p gpurun_out
: > gpurun_out/ring_ab2.txt
for cfg in "0 0" "128 0" "64 8" "128 8" "32 8" "0 8"; do
  set -- $cfg
  ABCOCT_RING_MB=$1 ABCOCT_HINTS=$2 timeout 200 python bench.py --workload c5-2048 --steps 10 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ring_mb', $1, 'hints', $2, 'value %.4e' % d['value'], 'frac %.4f' % d['roofline']['frac'], 'ok', d['e2e']['matches_device_leg'])
" >> gpurun_out/ring_ab2.txt
done
for cfg in "64 8" "128 0"; do
  set -- $cfg
  ABCOCT_RING_MB=$1 ABCOCT_HINTS=$2 timeout 300 python bench.py --workload c5-2048 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ring_mb', $1, 'hints', $2, 'value %.4e' % d['value'], 'traffic', d['roofline'].get('traffic'))
" >> gpurun_out/ring_ab2.txt
done
cat gpurun_out/ring_ab2.txt
