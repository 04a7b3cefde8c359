mkdir -p gpurun_out/c6
timeout 900 python -m pytest tests/test_virtual_peer_gpu.py tests/test_self_peer_gpu.py -q --timeout 400 -p no:cacheprovider -k "bytegrad or moe or qadam" > gpurun_out/c6/pytest_sub.log 2>&1
tail -3 gpurun_out/c6/pytest_sub.log
for b in 32 64; do
BAGUA_BYTEGRAD_BLOCKS=$b timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --workloads bert > gpurun_out/c6/bert_b$b.json 2> gpurun_out/c6/bert_b$b.err
python -c "
import json; d=json.load(open('gpurun_out/c6/bert_b$b.json')); print('blocks=$b', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['gpu_launches'], d['config']['buckets'], d['clocks'])"
done
BAGUA_BYTEGRAD_MIN_BUCKET_BYTES=0 timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --workloads bert --no-e2e > gpurun_out/c6/bert_nomerge.json 2> gpurun_out/c6/bert_nomerge.err
python -c "
import json; d=json.load(open('gpurun_out/c6/bert_nomerge.json')); print('no merge', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'], d['config']['buckets'])"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --workloads bert --impl nccl_baseline --no-e2e > gpurun_out/c6/bert_nccl.json 2> gpurun_out/c6/bert_nccl.err
python -c "
import json; d=json.load(open('gpurun_out/c6/bert_nccl.json')); print('nccl arm', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'], d['config']['buckets'])"
timeout 300 python benchmarks/config_bench.py --config gpt2_moe --arm peer --steps 10 --warmup 3 > gpurun_out/c6/gpt2_moe_peer.json 2> gpurun_out/c6/gpt2_moe_peer.err; tail -c 420 gpurun_out/c6/gpt2_moe_peer.json
timeout 300 python scripts/kernel_zoo.py 2>/dev/null | grep -i "moe_\|bytegrad" | cut -c1-200
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --workloads vgg16 > gpurun_out/c6/vgg.json 2> gpurun_out/c6/vgg.err
python -c "
import json; d=json.load(open('gpurun_out/c6/vgg.json')); print('vgg', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['clocks'])"
