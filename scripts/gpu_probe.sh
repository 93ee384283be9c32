for pr in 0 1 2 4 6 8 9 3 15; do
  echo "== probe $pr"
  DADD_ATTN_PROBE=$pr timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 --dtype fp16 2>&1 | grep "^{" | head -2 | cut -c1-120
done
