mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/r02_topo_8gpu.txt
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_${N}gpu.json')); print($N, round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['h2d_ceiling']['pinned_h2d_gbs_this_rank'], d['e2e']['h2d_ceiling']['pinned_h2d_gbs_all_ranks'], 'numa', d['config'].get('numa_node_rank0'), 'u8', round(d['extras']['e2e_u8']['value']), 'cfg5', d['extras']['config5']['ms'], d['extras']['config5']['samples_per_s'], d['extras']['config5']['efficiency_vs_one_gpu_rate'], d['extras']['config5']['micro_batch'], d['clocks'])" || tail -20 gpurun_out/r02_bench_${N}gpu.err
done
head -14 gpurun_out/r02_topo_8gpu.txt; tail -6 gpurun_out/r02_topo_8gpu.txt
