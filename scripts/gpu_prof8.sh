mkdir -p gpurun_out
timeout 600 python scripts/profile_step.py --patients 8 > gpurun_out/profile_plain.log 2>&1 && ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_p8.csv python scripts/profile_step.py --patients 8 > gpurun_out/profile_ncu.log 2>&1; echo "rc=$?"
tail -n 2 gpurun_out/profile_plain.log
