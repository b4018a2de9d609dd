#!/bin/bash
# End-of-round evidence on one B200: GPU parity suite, default bench line (with cpu_baseline, in-run traffic), reference arm,
# per-workload lines, C5 batch sweep, ncu launch list + one full capture of the default fused kernel.
set -u
mkdir -p gpurun_out
T=r02_end
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -3 gpurun_out/${T}_pytest.log
python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err; echo "reference rc=$?"
: > gpurun_out/${T}_bench_all.jsonl
for wl in c5-2048 c5-1024 c5-4096 c1 c2 c3 c4; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-traffic --e2e-steps 2 2>> gpurun_out/${T}_bench_all.err >> gpurun_out/${T}_bench_all.jsonl
done
: > gpurun_out/${T}_batch_sweep.jsonl
for fr in 1 16 256 1024 4096; do
  python bench.py --workload c5-2048 --frames $fr --steps 10 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 2>> gpurun_out/${T}_bench_all.err >> gpurun_out/${T}_batch_sweep.jsonl
done
python - <<P | tee gpurun_out/${T}_bench_all.txt
import json
for fn in ('gpurun_out/${T}_bench_all.jsonl', 'gpurun_out/${T}_batch_sweep.jsonl'):
    for l in open(fn):
        d=json.loads(l); print(d['config']['name'], 'frames', d['config']['frames_per_step_per_gpu'], '%.3e A-scans/s'%d['value'], '%.3e B-scans/s'%d['bscans_per_s'], 'frac %.3f'%d['roofline']['frac'], 'regs', d['plan']['regs_per_thread'], 'kind', d['plan']['kernel_kind'], 'e2e %.3e'%d['e2e']['value'], 'ok', d['e2e']['matches_device_leg'])
P
python bench.py --steps 2 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 > gpurun_out/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 > gpurun_out/${T}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'wres_kernel|wrow_kernel|recon_kernel' -s 3 -c 1 -o /tmp/${T}_k -f python bench.py --steps 2 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 > gpurun_out/${T}_ncu2.log 2>&1
ncu -i /tmp/${T}_k.ncu-rep --page raw --csv > gpurun_out/${T}_kernel_raw.csv 2>/dev/null
python tools/ncu_sass_mix.py /tmp/${T}_k.ncu-rep > gpurun_out/${T}_kernel_sass_mix.txt 2>&1
python tools/ncu_by_line.py /tmp/${T}_k.ncu-rep 1048576 150 > gpurun_out/${T}_kernel_by_line.txt 2>&1
python tools/ncu_key.py gpurun_out/${T}_kernel_raw.csv
echo done
