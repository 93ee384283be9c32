# kernel micro-bench + ncu full captures of the top hot-path kernels (one GPU)
mkdir -p gpurun_out
timeout 600 python scripts/kbench.py > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"
cat gpurun_out/kbench.log
A="--kernel self_attn --iters 1 --no-flush"
timeout 300 python scripts/kbench.py $A > gpurun_out/kb_attn_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:self_attn_tc -s 3 -c 1 -f -o gpurun_out/attn_tc python scripts/kbench.py $A > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn rc=$?"
G="--kernel gn --iters 1 --no-flush"
timeout 300 python scripts/kbench.py $G > gpurun_out/kb_gn_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_nhwc -s 3 -c 1 -f -o gpurun_out/gn_nhwc python scripts/kbench.py $G > gpurun_out/ncu_gn.log 2>&1
echo "ncu gn rc=$?"
tail -3 gpurun_out/ncu_attn.log gpurun_out/ncu_gn.log
