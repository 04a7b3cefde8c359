mkdir -p gpurun_out/c5
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 5 --workloads bert --no-e2e --profile gpurun_out/c5/prof_ours > /dev/null 2> gpurun_out/c5/prof_ours.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 5 --workloads bert --no-e2e --impl ddp --profile gpurun_out/c5/prof_ddp > /dev/null 2> gpurun_out/c5/prof_ddp.err
for v in 0 1; do
BAGUA_NATIVE_HOOKS=$v timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --workloads bert --no-e2e > gpurun_out/c5/bert_hooks$v.json 2> gpurun_out/c5/bert_hooks$v.err
python -c "
import json; d=json.load(open('gpurun_out/c5/bert_hooks$v.json')); print('native_hooks=$v', d['value'], d['ms_per_step'], d['gpu_launches'])"
done
ls gpurun_out/c5; head -45 gpurun_out/c5/prof_ours.bert.txt | cut -c1-200
