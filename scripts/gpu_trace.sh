mkdir -p gpurun_out
timeout 300 python scripts/attn_trace.py > gpurun_out/attn_trace.log 2>&1; echo "trace rc=$?"; head -90 gpurun_out/attn_trace.log
