#!/bin/bash
# Regenerates the round's evidence on a B200 (run under gpurun from the repo root):
#   launch list (per-kernel durations), one `ncu --set full` capture of one step, the bench lines.
set -x
R=${1:-r01}
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_pile100k.csv \
    python bench.py --steps 3 --warmup 3 --profile-range --no-cpu > gpurun_out/capture.log 2>&1
python profiles/summarize_launches.py gpurun_out/${R}_launches_pile100k.csv > gpurun_out/${R}_launches_pile100k.txt
ncu --profile-from-start off --set full --clock-control none --import-source on -c 24 -f -o gpurun_out/${R}_full_pile100k \
    python bench.py --steps 1 --warmup 3 --profile-range --no-cpu >> gpurun_out/capture.log 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_batch4096.csv \
    python profiles/batch_ncu.py >> gpurun_out/capture.log 2>&1
python profiles/summarize_launches.py gpurun_out/${R}_launches_batch4096.csv > gpurun_out/${R}_launches_batch4096.txt
python bench.py > gpurun_out/${R}_bench_n1_pile100k.json 2>> gpurun_out/capture.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_n1_reference_arm.json 2>> gpurun_out/capture.log
tail -3 gpurun_out/capture.log
