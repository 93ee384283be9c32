# usage: gpurun --gpus N -- 'bash scripts/gpu_scale.sh N'
N=${1:-2}
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/r2/bench_n$N.json 2> gpurun_out/r2/bench_n$N.err; echo "bench n$N rc=$?"
python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2/bench_n$N.json') if l.startswith('{')][-1]); print('lines in file:', sum(1 for _ in open('gpurun_out/r2/bench_n$N.json'))); print('N=$N value', d['value'], 'e2e', d['e2e']['value'], 'ms/step', d['ms_per_step'], d['clocks']); print('strong', d.get('strong')); print('config3', {k: v for k, v in d['config3'].items() if k != 'workload'})"
tail -2 gpurun_out/r2/bench_n$N.err
