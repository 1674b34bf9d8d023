N=$1
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --steps 3 --warmup 3 > $O/r02_final_bench_${N}gpu.json 2> $O/r02_final_bench_${N}gpu.err
$TR bench.py --gpus $N --scaling strong --total-games 16384 --steps 2 --warmup 2 > $O/r02_final_bench_${N}gpu_strong16k.json 2> $O/r02_final_bench_${N}gpu_strong16k.err
for f in $O/r02_final_bench_${N}gpu.json $O/r02_final_bench_${N}gpu_strong16k.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', round(d['ms_per_step'],1), 'wall ms', round(d['e2e']['wall_ms_per_step'],1))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
