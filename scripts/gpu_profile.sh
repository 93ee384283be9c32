mkdir -p gpurun_out
P=${1:-2}; DT=${2:-bf16}
python scripts/profile_step.py --patients $P --dtype $DT > gpurun_out/profile_plain.log 2>&1 && \
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_p${P}_${DT}.csv python scripts/profile_step.py --patients $P --dtype $DT > gpurun_out/profile_ncu.log 2>&1
echo "rc=$?"; cat gpurun_out/profile_plain.log | tail -3; tail -3 gpurun_out/profile_ncu.log; wc -l gpurun_out/launches_p${P}_${DT}.csv
