#!/bin/bash
# Run every bench workload once (device-resident + e2e), one JSON line each, into gpurun_out/bench_all.jsonl
out=gpurun_out/bench_all.jsonl
: > $out
for wl in c5-2048 c5-1024 c5-4096 c1 c2 c4 c3; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --e2e-steps 2 2>>gpurun_out/bench_all.err | tail -1 >> $out
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_all.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad line', l[:100]); continue
    r=d['roofline']
    print(f"{d['config']['name']:8s} value {d['value']:.3e} A-scans/s  {d['bscans_per_s']:.0f} B-scans/s  ms/step {d['ms_per_step']:.3f}  recon {r['achieved']:.0f} GB/s = {100*r['frac']:.1f}%  e2e {d['e2e']['value']:.3e}  launches {d['gpu_launches']}  G={d['plan']['groups_per_cta']} regs={d['plan']['regs_per_thread']}")
PY
