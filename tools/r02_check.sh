#!/bin/bash
# Round-2 checkpoint on one B200: GPU parity suite, default bench line, per-workload lines (new kernel and ABCOCT_KERNEL=1 = the
# round-1 group kernel), ncu launch list and one full capture of the fused kernel summarised on the box.
set -u
mkdir -p gpurun_out
T=${1:-r02}
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"
: > gpurun_out/${T}_bench_all.jsonl
for wl in c5-2048 c5-1024 c5-4096 c1 c2 c3 c4; do
  for mode in new old; do
    [ $mode = old ] && env="ABCOCT_KERNEL=1" || env="ABCOCT_KERNEL=0"
    env $env python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-traffic --e2e-steps 2 2>> gpurun_out/${T}_bench_all.err | sed "s/^{/{\"mode\": \"$mode\", /" >> gpurun_out/${T}_bench_all.jsonl
  done
done
python - <<P | tee gpurun_out/${T}_bench_all.txt
import json
for l in open('gpurun_out/${T}_bench_all.jsonl'):
    d=json.loads(l); print(d['config']['name'], d['mode'], '%.3e'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'regs', d['plan']['regs_per_thread'], 'e2e %.3e'%d['e2e']['value'], 'ok', d['e2e']['matches_device_leg'], d['roofline']['kernel'][:12])
P
python bench.py --steps 2 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'wres_kernel|wrow_kernel|recon_kernel' -s 3 -c 1 -o /tmp/${T}_wrow -f python bench.py --steps 2 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 > gpurun_out/${T}_ncu2.log 2>&1
ncu -i /tmp/${T}_wrow.ncu-rep --page raw --csv > gpurun_out/${T}_wrow_raw.csv 2>/dev/null
python tools/ncu_sass_mix.py /tmp/${T}_wrow.ncu-rep > gpurun_out/${T}_wrow_sass_mix.txt 2>&1
python tools/ncu_by_line.py /tmp/${T}_wrow.ncu-rep 1048576 150 > gpurun_out/${T}_wrow_by_line.txt 2>&1
cp /tmp/${T}_wrow.ncu-rep gpurun_out/ 2>/dev/null
echo done
