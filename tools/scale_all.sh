# scaling run of the default bench command at 1/2/4/8 GPUs of one box (what the driver does at round end)
for n in 8 4 2; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 40 --warmup 5 2>&1 | grep "^{" > gpurun_out/r01_bench_c5_n$n.json
  python -c "
import json; d=json.load(open('gpurun_out/r01_bench_c5_n$n.json')); print($n, d['value'], d['ms_per_step'], d['step_roofline']['frac'], d['e2e']['value'] if d['e2e'] else None)"
done
timeout 200 python bench.py --workload c5 --steps 40 --warmup 5 > gpurun_out/r01_bench_c5_n1.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r01_bench_c5_n1.json')); print(1, d['value'], d['ms_per_step'], d['step_roofline']['frac'], d['roofline']['achieved'], d['e2e']['value'])"
