#!/bin/bash
# One full ncu capture of the fused kernel per configuration given as "NAME ENV=VAL ..." lines on stdin.  The reports are large
# (35 MB each), so they are summarised on the GPU box (raw metrics csv, opcode mix, per-source-line view) and only the first
# report travels back.
set -u
mkdir -p gpurun_out
first=1
while read -r name envs; do
  [ -z "$name" ] && continue
  env $envs python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${name}_plain.log 2>&1 || { echo "plain run failed: $name"; continue; }
  grep -o '"value": [0-9.e+]*' gpurun_out/${name}_plain.log | head -1
  env $envs ncu --set full --clock-control none --import-source on -k regex:'wres_kernel|wrow_kernel|recon_kernel' -s 3 -c 1 -o /tmp/$name -f python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/${name}_ncu.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  python tools/ncu_sass_mix.py /tmp/$name.ncu-rep > gpurun_out/${name}_sass_mix.txt 2>&1
  python tools/ncu_by_line.py /tmp/$name.ncu-rep 1048576 120 > gpurun_out/${name}_by_line.txt 2>&1
  if [ $first = 1 ]; then cp /tmp/$name.ncu-rep gpurun_out/; first=0; fi
done
ls -la gpurun_out/ | head -30
