DBN_B200_DP_CHUNKS=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 30 --warmup 3 --no-e2e 2>&1 | grep "^{" > gpurun_out/n8_chunks1.json
python -c "
import json; d=json.load(open('gpurun_out/n8_chunks1.json')); print('chunks=1', d['value'], d['ms_per_step'])"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 30 --warmup 3 2>&1 | grep "^{" > gpurun_out/r01_bench_c5_n8.json
python -c "
import json; d=json.load(open('gpurun_out/r01_bench_c5_n8.json')); print('default', d['value'], d['ms_per_step'], d['e2e'], d['step_roofline']['frac'])"
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --impl reference --steps 2 --warmup 0 2>&1 | grep "^{" | cut -c1-200
