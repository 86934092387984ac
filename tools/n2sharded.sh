timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/dp_check.py 2>&1 | grep "dp_check\|Error\|error" | tail -6
run() { # sharded gb tag port
  env DBN_B200_DP_SHARDED=$1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 2 --global-batch $2 --steps 40 --warmup 5 --no-e2e 2>&1 | grep "^{" > gpurun_out/n2sh_$1_$2.json
  python -c "
import json; d=json.load(open('gpurun_out/n2sh_$1_$2.json')); print('sharded=$1 global_batch=$2', d['value'], d['ms_per_step'])"
}
run 1 32768 a 29552
run 0 32768 a 29553
run 1 8192 a 29554
run 0 8192 a 29555
