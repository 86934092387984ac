run() { # sharded port
  env DBN_B200_DP_SHARDED=$1 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 8 --steps 40 --warmup 5 --no-e2e 2>&1 | grep "^{" > gpurun_out/n8sh_$1.json
  python -c "
import json; d=json.load(open('gpurun_out/n8sh_$1.json')); print('sharded=$1', d['value'], d['ms_per_step'])"
}
run 1 29561
run 0 29562
