N=4
out=gpurun_out/g$N
mkdir -p $out
timeout 1500 python -m pytest tests/test_peer_gpu.py -q --timeout 600 -p no:cacheprovider > $out/pytest_peer.log 2>&1
echo "pytest rc=$?" >> $out/pytest_peer.log
tail -6 $out/pytest_peer.log
port=29700
for impl in ours nccl_baseline ddp; do
port=$((port+1))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --impl $impl > $out/bench_$impl.json 2> $out/bench_$impl.err
echo "bench $impl rc=$?" >> $out/bench_$impl.err
python - <<PY
import json
try:
    d=json.load(open("$out/bench_$impl.json"))
    b=d.get("bert_large_bytegrad",{})
    print("$impl N=$N", "vgg", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") and "value" in d["e2e"] else d.get("e2e"), "launches", d["gpu_launches"], "bert", round(b.get("value",0),2), "e2e", (b.get("e2e") or {}).get("value"), d.get("verify",{}).get("max_diff_in_ulps_of_the_weight_dtype"), d["config"].get("allreduce_variants"), b.get("config",{}).get("allreduce_variants"), d["clocks"])
except Exception as e:
    print("$impl failed", e)
PY
done
# autotune through the elastic launcher (reference CI: .buildkite/scripts/benchmark.sh:14-37), then the static default for comparison
timeout 300 python -m bagua_b200.distributed.run --standalone --nnodes=1 --nproc_per_node $N --autotune_level 1 --is_output_autotune_log --autotune_warmup_time 2 --autotune_max_samples 12 --autotune_sampling_confidence_time 1 examples/benchmark/synthetic_benchmark.py --num-iters 160 --model vgg16 --bf16 > $out/autotune_vgg16.log 2>&1
echo "autotune rc=$?"; grep -i "img/sec per GPU:\|Total img\|autotune\|bucket" $out/autotune_vgg16.log | tail -25
timeout 200 python -m bagua_b200.distributed.run --standalone --nnodes=1 --nproc_per_node $N examples/benchmark/synthetic_benchmark.py --num-iters 30 --model vgg16 --bf16 > $out/static_vgg16.log 2>&1
echo "static rc=$?"; grep "Img/sec per GPU:\|Total img" $out/static_vgg16.log
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port benchmarks/collective_bench.py --out $out/allreduce_n$N.json --iters 10 --blocks 8,16,32 --sizes 1048576,16777216,268435456 > $out/collective.log 2>&1
tail -24 $out/collective.log
