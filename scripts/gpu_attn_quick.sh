mkdir -p gpurun_out/r2
echo "== kbench B=104"
timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 --dtype fp16 2>&1 | grep "^{\|timeout rec\|Error" | cut -c1-200
echo "== kbench bf16 B=104"
timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 --dtype bf16 2>&1 | grep "^{\|timeout rec\|Error" | head -2 | cut -c1-200
echo "== kbench B=13"
timeout 300 python scripts/kbench.py --kernel self_attn --batch 13 --dtype fp16 2>&1 | grep "^{\|timeout rec\|Error" | cut -c1-200
echo "== kbench res64 B=16"
timeout 300 python scripts/kbench.py --kernel self_attn --batch 16 --res 64 --dtype fp16 2>&1 | grep "^{\|timeout rec\|Error" | cut -c1-200

echo "== trace"
DADD_ATTN_TRACE=gpurun_out/r2/attn_trace.txt timeout 120 python scripts/kbench.py --kernel self_attn --batch 26 --iters 1 --no-flush --dtype fp16 2>&1 | tail -1
python scripts/attn_trace.py gpurun_out/r2/attn_trace.txt | grep "^warp 17\|^warp  0\|^warp  8"
