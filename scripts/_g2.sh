N=${1:-2}
out=gpurun_out/g$N
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/smi.txt
timeout 1500 python -m pytest tests/test_peer_gpu.py -q --timeout 600 -p no:cacheprovider > $out/pytest_peer.log 2>&1
echo "pytest rc=$?" >> $out/pytest_peer.log
tail -8 $out/pytest_peer.log
port=29600
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
port=$((port+1))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port benchmarks/collective_bench.py --out $out/allreduce_n$N.json --iters 10 --blocks 8,16,32,64 > $out/collective.log 2>&1
tail -30 $out/collective.log
