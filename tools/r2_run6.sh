mkdir -p gpurun_out
timeout 600 python tools/gate_survey.py > gpurun_out/r2_gate_survey.txt 2>&1; cat gpurun_out/r2_gate_survey.txt | grep -v Warning
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench6_2gpu.json 2> gpurun_out/r2_bench6_2gpu.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench6_2gpu.json')); print(round(d['value']), round(d['e2e']['value']), d['e2e']['h2d_ceiling'], d['config'].get('numa_node_rank0'), d['extras'].get('config5'), round(d['extras']['e2e_u8']['value']))" || tail -20 gpurun_out/r2_bench6_2gpu.err
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1; head -12 gpurun_out/r2_topo.txt; numactl --hardware 2>/dev/null | head -5; lscpu | grep -i "numa\|socket\|model name" | head
