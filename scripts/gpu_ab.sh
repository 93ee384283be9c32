run() { timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 --dtype fp16 2>&1 | grep "^{" | head -2 | cut -c1-140; }
cp progressive_stable_diffusion_b200/libdadd_b200.so /tmp/orig.so
for v in A B; do
  cp scripts/micro/variants/lib$v.so progressive_stable_diffusion_b200/libdadd_b200.so
  echo "== variant $v (A = epilogue at item end, B = deferred), KS"; run
  echo "== variant $v, runtime K loop"; DADD_ATTN_NOKS=1 run
done
cp /tmp/orig.so progressive_stable_diffusion_b200/libdadd_b200.so
