#!/bin/bash
# dB scratch as a small reused region (ABCOCT_SCRATCH_MB -> B-scans per launch) instead of one region per B-scan of the batch:
# throughput and DRAM traffic of the fused kernel on C5-2048
set -u
mkdir -p gpurun_out
: > gpurun_out/scratch_sweep.txt
for mb in 8192 1024 256 128 64 32; do
  ABCOCT_SCRATCH_MB=$mb python bench.py --workload c5-2048 --steps 10 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('scratch_mb', $mb, 'value %.4e' % d['value'], 'launches', d['gpu_launches'], 'traffic', d['roofline'].get('traffic'), 'e2e %.3e' % d['e2e']['value'])
" >> gpurun_out/scratch_sweep.txt
done
cat gpurun_out/scratch_sweep.txt
