run() { # reserve chunks extra_env
  env DBN_B200_SM_RESERVE=$1 DBN_B200_DP_CHUNKS=$2 $3 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $4 bench.py --gpus 8 --steps 30 --warmup 3 --no-e2e 2>&1 | grep "^{" > gpurun_out/n8_$1_$2_$5.json
  python -c "
import json; d=json.load(open('gpurun_out/n8_$1_$2_$5.json')); print('reserve=$1 chunks=$2 $3', d['value'], d['ms_per_step'])"
}
run 0 4 FOO=1 29521 a
run 16 4 FOO=1 29522 a
run 32 4 FOO=1 29523 a
run 16 2 FOO=1 29524 a
run 16 8 FOO=1 29525 a
run 16 4 NCCL_MAX_CTAS=8 29526 b
