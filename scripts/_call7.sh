mkdir -p gpurun_out/c7
run() { name=$1; shift; timeout 300 env "$@" python bench.py --gpus 1 --steps 20 --warmup 5 --workloads bert --no-e2e > gpurun_out/c7/$name.json 2> gpurun_out/c7/$name.err; python -c "
import json
try:
    d=json.load(open('gpurun_out/c7/$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],2), d['gpu_launches'], d['config']['buckets'])
except Exception as e: print('$name failed', e)"; }
run coop0_merge_b64 BAGUA_COOPERATIVE_LAUNCH=0 BAGUA_BYTEGRAD_BLOCKS=64
run coop0_merge_b128 BAGUA_COOPERATIVE_LAUNCH=0 BAGUA_BYTEGRAD_BLOCKS=128
run coop1_merge_b128 BAGUA_BYTEGRAD_BLOCKS=128
run coop0_nomerge_b32 BAGUA_COOPERATIVE_LAUNCH=0 BAGUA_BYTEGRAD_MIN_BUCKET_BYTES=0
run coop0_nomerge_b64 BAGUA_COOPERATIVE_LAUNCH=0 BAGUA_BYTEGRAD_MIN_BUCKET_BYTES=0 BAGUA_BYTEGRAD_BLOCKS=64
run coop0_merge32M_b64 BAGUA_COOPERATIVE_LAUNCH=0 BAGUA_BYTEGRAD_MIN_BUCKET_BYTES=33554432 BAGUA_BYTEGRAD_BLOCKS=64
timeout 300 python benchmarks/config_bench.py --config gpt2_moe --arm peer --steps 10 --warmup 3 > gpurun_out/c7/gpt2_moe_peer.json 2> gpurun_out/c7/gpt2_moe_peer.err; tail -c 330 gpurun_out/c7/gpt2_moe_peer.json; grep -c Error gpurun_out/c7/gpt2_moe_peer.err
