run() { # chunks nocomm tag
  env DBN_B200_SM_RESERVE=0 DBN_B200_DP_CHUNKS=$1 DBN_B200_DP_NOCOMM=$2 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --global-batch 8192 --steps 40 --warmup 5 --no-e2e 2>&1 | grep "^{" > gpurun_out/n2emu_$1_$2.json
  python -c "
import json; d=json.load(open('gpurun_out/n2emu_$1_$2.json')); print('chunks=$1 nocomm=$2', d['value'], d['ms_per_step'])"
}
run 1 0 29531
run 1 1 29532
run 2 0 29533
run 2 1 29534
run 4 0 29535
run 4 1 29536
run 8 1 29537
