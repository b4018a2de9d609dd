#!/bin/bash
# ncu capture of the two kernels behind C3 (rowprep_kernel with the Fourier upsample, recon_kernel IN_F32 on N = 3840)
set -u
mkdir -p gpurun_out
python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 > gpurun_out/c3_plain.json 2> gpurun_out/c3_plain.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:'rowprep_kernel' -s 3 -c 1 -o /tmp/c3_prep -f python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --no-traffic --e2e-steps 1 > gpurun_out/c3_ncu_prep.log 2>&1
ncu -i /tmp/c3_prep.ncu-rep --page raw --csv > gpurun_out/c3_prep_raw.csv 2>/dev/null
python tools/ncu_key.py gpurun_out/c3_prep_raw.csv > gpurun_out/c3_prep_key.txt 2>&1
python tools/ncu_sass_mix.py /tmp/c3_prep.ncu-rep > gpurun_out/c3_prep_sass_mix.txt 2>&1
python tools/ncu_by_line.py /tmp/c3_prep.ncu-rep 2400 60 > gpurun_out/c3_prep_by_line.txt 2>&1
cat gpurun_out/c3_prep_key.txt
