mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -p no:cacheprovider -k "self_attention" 2>&1 | tail -1
timeout 300 python scripts/kbench.py --kernel self_attn --batch 104 2>&1 | head -2
timeout 300 python scripts/kbench.py --kernel self_attn --batch 26 2>&1 | head -2
