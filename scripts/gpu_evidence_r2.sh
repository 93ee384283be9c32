# Round-2 evidence: launch list of one denoising step of the bench + ncu --set full captures (cold caches) of the hot kernels
mkdir -p gpurun_out/r2/ncu
O=gpurun_out/r2/ncu
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench_step_B104.csv python bench.py --steps 1 --warmup 3 --profile-step --no-extras > $O/bench_under_ncu.log 2>&1; echo "ncu launch list rc=$?"
cap() {  # name, kernel regex, kbench kernel, extra kbench args
  timeout 600 ncu --set full --import-source on --cache-control all --clock-control none -k "regex:$2" -c 1 -f -o $O/$1 python scripts/kbench.py --kernel $3 --batch 104 --iters 1 --no-flush --dtype fp16 $4 > $O/ncu_$1.log 2>&1; echo "ncu $1 rc=$?"
  python scripts/ncu_pick.py $O/$1.ncu-rep > $O/$1.txt 2>&1
}
cap self_attn_tc self_attn_tc self_attn
cap gn_cluster gn_cluster gn
cap cross_attn_tc cross_attn_tc cross
cap add_layernorm layernorm_kernel add_ln
ls -la $O
# keep only the attention report (size limit of the pull-back); the others are summarised in the .txt files
rm -f $O/gn_cluster.ncu-rep $O/cross_attn_tc.ncu-rep $O/add_layernorm.ncu-rep
