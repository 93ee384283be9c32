mkdir -p gpurun_out
for MB in 0 32 48 96; do for P in 2 8; do
echo "FF_SLICE_MB=$MB P=$P: $(DADD_FF_SLICE_MB=$MB timeout 600 python scripts/profile_step.py --patients $P 2>&1 | grep 'graph replay')"
done; done
