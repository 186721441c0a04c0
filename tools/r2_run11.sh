mkdir -p gpurun_out
for N in 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_${N}gpu.json')); print($N, round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e'].get('frac_of_h2d_ceiling'), d['e2e']['h2d_ceiling']['pinned_h2d_gbs_this_rank'], d['e2e']['h2d_ceiling']['pinned_h2d_gbs_all_ranks'], 'numa', d['config'].get('numa_node_rank0'), 'u8', round(d['extras']['e2e_u8']['value']), 'cfg5', d['extras']['config5']['ms'], d['extras']['config5']['samples_per_s'], d['extras']['config5']['efficiency_vs_one_gpu_rate'], d['extras']['config5']['micro_batch'], d['clocks'])" || grep -v "^W\|Warn" gpurun_out/r02_bench_${N}gpu.err | tail -20
done
timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu -x 2>&1 | tail -4
