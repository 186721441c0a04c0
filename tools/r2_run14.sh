mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu_taper.json 2> gpurun_out/r02_bench_${N}gpu_taper.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_${N}gpu_taper.json')); print($N, round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('copy_share_of_call'), d['e2e'].get('frac_of_h2d_ceiling'), d['e2e']['h2d_ceiling'], 'u8', round(d['extras']['e2e_u8']['value']), 'cfg5', d['extras']['config5']['samples_per_s'], d['clocks'])" || grep -v "^W\|Warn" gpurun_out/r02_bench_${N}gpu_taper.err | tail -20
