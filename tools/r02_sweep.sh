#!/bin/bash
# usage: r02_sweep.sh < config lines "NAME WORKLOAD ENV=VAL ..." ; per config: a plain bench run (value) and a light ncu pass
# (kernel time, DRAM bytes, L2 hit rate, issue utilisation, instructions, local-memory instructions)
set -u
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed
: > gpurun_out/sweep.txt
while read -r name wl envs; do
  [ -z "$name" ] && continue
  v=$(env $envs python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --e2e-steps 1 2>gpurun_out/sweep_${name}.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4e frac %.3f regs %d ok %s'%(d['value'],d['roofline']['frac'],d['plan']['regs_per_thread'],d['e2e']['matches_device_leg']))")
  env $envs ncu --metrics $M --clock-control none -k regex:'wres_kernel|wrow_kernel|recon_kernel' -s 3 -c 1 --csv python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu --e2e-steps 1 2>/dev/null | python -c "
import sys,csv
rows=[r for r in csv.reader(sys.stdin) if len(r)>10 and r[0].isdigit()]
print(' '.join('%s=%s'%(r[-3].split('__')[-1][:28],r[-1]) for r in rows))" > gpurun_out/sweep_${name}.ncu 2>&1
  echo "$name $wl | $v | $(cat gpurun_out/sweep_${name}.ncu)" | tee -a gpurun_out/sweep.txt
done
