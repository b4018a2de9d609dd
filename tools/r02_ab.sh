#!/bin/bash
# Round-2 A/B on one B200: GPU parity suite, then bench of the warp-per-A-scan kernel (16 and 12 warps per CTA) against the
# group-per-row-pair kernel (ABCOCT_KERNEL=1) on every single-GPU workload, then the ncu launch list and one full capture.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -3 gpurun_out/r02_pytest.log
: > gpurun_out/r02_ab.jsonl
for wl in c5-2048 c1 c2 c4 c5-1024; do
  for mode in "w16" "w12" "old"; do
    case $mode in
      w16) env="ABCOCT_WROW_NW=16";;
      w12) env="ABCOCT_WROW_NW=12";;
      old) env="ABCOCT_KERNEL=1";;
    esac
    echo "== $wl $mode" >> gpurun_out/r02_ab.err
    env $env python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --e2e-steps 1 2>> gpurun_out/r02_ab.err | sed "s/^{/{\"mode\": \"$mode\", /" >> gpurun_out/r02_ab.jsonl
  done
done
python - <<'P'
import json
for l in open('gpurun_out/r02_ab.jsonl'):
    d=json.loads(l); print(d['config']['name'], d['mode'], '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'regs', d['plan']['regs_per_thread'], 'ok', d['e2e']['matches_device_leg'])
P
python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r02_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wrow_kernel -s 3 -c 1 -o gpurun_out/r02_wrow_v0 -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r02_ncu2.log 2>&1
echo done
