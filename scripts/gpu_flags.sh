for f in 0 1 2 3; do
  echo "== flags $f"
  DADD_ATTN_FLAGS=$f timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 --dtype fp16 2>&1 | grep "^{" | head -2 | cut -c1-150
done
