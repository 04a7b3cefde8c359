mkdir -p gpurun_out/c9
timeout 900 python -m pytest tests/test_virtual_peer_gpu.py tests/test_self_peer_gpu.py -q --timeout 400 -p no:cacheprovider -k "bytegrad or qadam or moe" > gpurun_out/c9/pytest_sub.log 2>&1
tail -4 gpurun_out/c9/pytest_sub.log
timeout 300 python scripts/kernel_zoo.py 2>/dev/null | grep -i "bytegrad" | cut -c1-200
for b in 32 64; do
BAGUA_BYTEGRAD_BLOCKS=$b timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --workloads bert --no-e2e > gpurun_out/c9/bert_b$b.json 2> gpurun_out/c9/bert_b$b.err
python -c "
import json
try:
    d=json.load(open('gpurun_out/c9/bert_b$b.json')); print('blocks=$b', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'], d['config']['buckets'])
except Exception as e: print('failed', e)"
done
BAGUA_BYTEGRAD_MIN_BUCKET_BYTES=0 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --workloads bert --no-e2e > gpurun_out/c9/bert_nomerge.json 2> gpurun_out/c9/bert_nomerge.err
python -c "
import json; d=json.load(open('gpurun_out/c9/bert_nomerge.json')); print('no merge', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'], d['config']['buckets'])"
timeout 300 python benchmarks/config_bench.py --config gpt2_moe --arm peer --steps 10 --warmup 3 > gpurun_out/c9/gpt2_moe_peer.json 2> gpurun_out/c9/gpt2_moe_peer.err; tail -c 330 gpurun_out/c9/gpt2_moe_peer.json | head -c 120; grep -c Error gpurun_out/c9/gpt2_moe_peer.err
