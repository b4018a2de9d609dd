#!/bin/bash
# A/B of the dB scratch ring of wrow_kernel (ABCOCT_RING_MB; 0 = one region per B-scan) on C5-2048: throughput, then DRAM traffic.
# Needs a library built with ABCOCT_BUILD_RING=1 (-DABC_WROW_RING); the product build ignores ABCOCT_RING_MB.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "scratch_ring or test_against_oracle or full_size or golden" 2>&1 | tail -3
: > gpurun_out/ring_ab.txt
for mb in 0 64 32 128 256 0 64; do
  ABCOCT_RING_MB=$mb timeout 200 python bench.py --workload c5-2048 --steps 10 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ring_mb', $mb, 'value %.4e' % d['value'], 'frac %.4f' % d['roofline']['frac'], 'e2e %.3e' % d['e2e']['value'], 'ok', d['e2e']['matches_device_leg'])
" >> gpurun_out/ring_ab.txt
done
for mb in 64; do
  ABCOCT_RING_MB=$mb timeout 300 python bench.py --workload c5-2048 --steps 5 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ring_mb', $mb, 'value %.4e' % d['value'], 'traffic', d['roofline'].get('traffic'), 'algorithmic', d['roofline']['algorithmic_bytes_per_launch'])
" >> gpurun_out/ring_ab.txt
done
cat gpurun_out/ring_ab.txt
