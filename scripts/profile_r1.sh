#!/usr/bin/env bash
# Run under gpurun (one GPU).  ncu launch list + one --set full capture of K3 and K1 for the default
# bench command, plus the full-size parity report.  Outputs land in gpurun_out/.
set -uo pipefail
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err   # builds + caches the index
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 > gpurun_out/launches_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hnsw_walk -s 1 -c 1 -f -o gpurun_out/prof_k3 \
    python bench.py --steps 2 --warmup 1 > gpurun_out/prof_k3.log 2>&1
python scripts/parity_c2.py > gpurun_out/parity_c2.json 2> gpurun_out/parity_c2.err
python scripts/parity_c2.py --dist blobs > gpurun_out/parity_c2_blobs.json 2> gpurun_out/parity_c2_blobs.err
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_r1.json 2> gpurun_out/bench_ref_r1.err
ls -la gpurun_out
