N=8
out=gpurun_out/g$N
mkdir -p $out
port=29800
for impl in ours nccl_baseline ddp; do
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --impl $impl > $out/bench_$impl.json 2> $out/bench_$impl.err
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
for arm in peer nccl; do
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port benchmarks/config_bench.py --config gpt2_moe --arm $arm --steps 10 --warmup 3 > $out/gpt2_moe_$arm.json 2> $out/gpt2_moe_$arm.err
echo "moe $arm rc=$?"; tail -c 700 $out/gpt2_moe_$arm.json | cut -c1-330
done
timeout 200 python -m pytest tests/test_peer_gpu.py -q --timeout 180 -p no:cacheprovider -k "hierarchical" > $out/pytest_hier.log 2>&1; tail -3 $out/pytest_hier.log
