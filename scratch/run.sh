set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python profiles/step_breakdown.py pile100k 2>&1 | tail -8
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_new.csv python bench.py --steps 3 --warmup 3 --profile-range --no-cpu > gpurun_out/b.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_new.csv | head -30
