#!/bin/bash
# usage: r02_vals.sh < lines "NAME WORKLOAD ENV=VAL ..." ; one plain bench run per line (device-resident value only, no ncu)
set -u
mkdir -p gpurun_out
: > gpurun_out/vals.txt
while read -r name wl envs; do
  [ -z "$name" ] && continue
  v=$(env $envs python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 2>gpurun_out/vals_${name}.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4e frac %.3f regs %d kind %d nw %d ok %s'%(d['value'],d['roofline']['frac'],d['plan']['regs_per_thread'],d['plan'].get('kernel_kind',-1),d['plan']['groups_per_cta'],d['e2e']['matches_device_leg']))")
  echo "$name $wl | $v" | tee -a gpurun_out/vals.txt
done
