mkdir -p gpurun_out; rm -f gpurun_out/kb_gn_flat.log
for B in 26 104; do
DADD_GN_FLAT=1 timeout 300 python scripts/kbench.py --kernel gn --batch $B >> gpurun_out/kb_gn_flat.log 2>&1
done
cat gpurun_out/kb_gn_flat.log
