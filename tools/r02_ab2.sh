#!/bin/bash
# Round-2 A/B #2: kernel v1 (predicate-free pre-phase, register prefetch of the next row, late publish, non-inlined normalise)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -3 gpurun_out/r02b_pytest.log
: > gpurun_out/r02b_ab.jsonl
run() { # name workload env...
  local name=$1 wl=$2; shift 2
  echo "== $wl $name" >> gpurun_out/r02b_ab.err
  env "$@" python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --e2e-steps 1 2>> gpurun_out/r02b_ab.err | sed "s/^{/{\"mode\": \"$name\", /" >> gpurun_out/r02b_ab.jsonl
}
for ns in 1 8 32; do for nw in 16 12; do run "w${nw}_ns${ns}" c5-2048 ABCOCT_WROW_NW=$nw ABCOCT_NSPLIT=$ns; done; done
for wl in c1 c2 c4 c5-1024; do for nw in 16 12; do run "w${nw}" $wl ABCOCT_WROW_NW=$nw; done; done
python - <<'P'
import json
for l in open('gpurun_out/r02b_ab.jsonl'):
    d=json.loads(l); print(d['config']['name'], d['mode'], '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'regs', d['plan']['regs_per_thread'], 'ok', d['e2e']['matches_device_leg'])
P
for ns in 1 8; do
ABCOCT_NSPLIT=$ns ncu --set full --clock-control none --import-source on -k regex:wrow_kernel -s 3 -c 1 -o gpurun_out/r02b_wrow_ns$ns -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r02b_ncu_ns$ns.log 2>&1
done
echo done
