timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_batch.csv python profiles/batch_ncu.py > gpurun_out/b.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_batch.csv | head -30
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_new.csv python bench.py --steps 3 --warmup 3 --profile-range --no-cpu > gpurun_out/b.log 2>&1
python profiles/summarize_launches.py gpurun_out/launches_new.csv | head -30
