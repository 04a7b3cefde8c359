mkdir -p gpurun_out/c8
for impl in ours ddp; do for bs in 6 24; do
timeout 300 python bench.py --gpus 1 --steps 15 --warmup 5 --workloads bert --no-e2e --impl $impl --bert-batch-size $bs > gpurun_out/c8/${impl}_bs$bs.json 2> gpurun_out/c8/${impl}_bs$bs.err
python -c "
import json
try:
    d=json.load(open('gpurun_out/c8/${impl}_bs$bs.json')); print('$impl bs$bs', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'host issue ms', round(d['host_issue_ms_per_step'] or 0,2), d['gpu_launches'])
except Exception as e: print('$impl $bs failed', e)"
done; done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 5 --workloads bert --no-e2e --profile gpurun_out/c8/prof_ours > /dev/null 2> gpurun_out/c8/prof_ours.err
grep -n "bytegrad_kernel\|flat_adam\|Self CUDA time total\|Self CPU time total\|aten::add_ \|AccumulateGrad  " gpurun_out/c8/prof_ours.bert.txt | cut -c1-60,90-260
