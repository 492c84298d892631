#!/bin/bash
# Regenerates round 2's evidence on a B200 (run under gpurun from the repo root): launch lists (per-kernel durations),
# `ncu --set full` captures of one step of pile100k and of batch4096x256, the SASS summary, the sanitizer runs.
set -x
R=r02
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_pile100k.csv \
    python bench.py --steps 3 --warmup 3 --profile-range --no-cpu --no-configs --batch-worlds 0 > gpurun_out/capture.log 2>&1
python profiles/summarize_launches.py gpurun_out/${R}_launches_pile100k.csv > gpurun_out/${R}_launches_pile100k.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_batch4096.csv \
    python profiles/batch_ncu.py >> gpurun_out/capture.log 2>&1
python profiles/summarize_launches.py gpurun_out/${R}_launches_batch4096.csv > gpurun_out/${R}_launches_batch4096.txt
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_batch512.csv \
    python profiles/batch_ncu.py 512 >> gpurun_out/capture.log 2>&1
python profiles/summarize_launches.py gpurun_out/${R}_launches_batch512.csv > gpurun_out/${R}_launches_batch512.txt
ncu --profile-from-start off --set full --clock-control none --import-source on -c 12 -f -o gpurun_out/${R}_full_pile100k \
    python bench.py --steps 1 --warmup 3 --profile-range --no-cpu --no-configs --batch-worlds 0 >> gpurun_out/capture.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -c 4 -f -o gpurun_out/${R}_full_batch4096 \
    python profiles/batch_ncu.py >> gpurun_out/capture.log 2>&1
for rep in pile100k batch4096; do
  ncu -i gpurun_out/${R}_full_${rep}.ncu-rep --page raw --csv > gpurun_out/${R}_raw_${rep}.csv 2>/dev/null
done
compute-sanitizer --tool memcheck --error-exitcode 0 python profiles/sanitize_probe.py > gpurun_out/${R}_memcheck.txt 2>&1
compute-sanitizer --tool racecheck --error-exitcode 0 python profiles/sanitize_probe.py > gpurun_out/${R}_racecheck.txt 2>&1
tail -3 gpurun_out/${R}_memcheck.txt gpurun_out/${R}_racecheck.txt
tail -3 gpurun_out/capture.log
