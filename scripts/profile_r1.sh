#!/usr/bin/env bash
# Run under gpurun (one GPU).  Produces the ncu launch list and one --set full capture of the
# dominant kernel (K3 walk) for the default bench command; outputs land in gpurun_out/.
set -uo pipefail
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err   # builds + caches the index
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hnsw_walk -s 1 -c 1 -f -o gpurun_out/prof_k3 \
    python bench.py --steps 2 --warmup 1 > gpurun_out/prof_k3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:adc_table -s 1 -c 1 -f -o gpurun_out/prof_k1 \
    python bench.py --steps 2 --warmup 1 > gpurun_out/prof_k1.log 2>&1
ls -la gpurun_out
