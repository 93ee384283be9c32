mkdir -p gpurun_out
A="--kernel self_attn --iters 1 --no-flush"
timeout 300 python scripts/kbench.py $A > gpurun_out/kb_attn_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:self_attn_tc2 -s 3 -c 1 -f -o gpurun_out/attn_tc2 python scripts/kbench.py $A > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"; tail -n 3 gpurun_out/ncu_attn.log
