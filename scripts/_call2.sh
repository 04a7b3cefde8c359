mkdir -p gpurun_out/c2
timeout 900 python -m pytest tests/test_self_peer_gpu.py tests/test_virtual_peer_gpu.py -q --timeout 400 -p no:cacheprovider -k "qadam or rebucket or fused_adam or combine_epilogue or graphed or allreduce_adam or weight_gate or sharded" > gpurun_out/c2/pytest_sub.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c2/pytest_sub.log
for impl in ours nccl_baseline ddp; do
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --impl $impl > gpurun_out/c2/bench_n1_$impl.json 2> gpurun_out/c2/bench_n1_$impl.err
echo "bench $impl rc=$?" >> gpurun_out/c2/bench_n1_$impl.err
done
timeout 400 python scripts/kernel_zoo.py > gpurun_out/c2/zoo_events.jsonl 2> gpurun_out/c2/zoo_events.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bagua -o gpurun_out/c2/ncu_zoo python scripts/kernel_zoo.py > gpurun_out/c2/ncu_zoo.log 2>&1
echo "ncu rc=$?" >> gpurun_out/c2/ncu_zoo.log
ls -la gpurun_out/c2
tail -4 gpurun_out/c2/pytest_sub.log
for impl in ours nccl_baseline ddp; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c2/bench_n1_$impl.json"))
    b=d.get("bert_large_bytegrad",{})
    print("$impl", "vgg", round(d["value"],1), "e2e", round(d["e2e"]["value"],1) if d.get("e2e") and "value" in d["e2e"] else d.get("e2e"), "launches", d["gpu_launches"], "bert", round(b.get("value",0),2), "e2e", (b.get("e2e") or {}).get("value"), d.get("verify"), d["config"].get("allreduce_variants"), b.get("config",{}).get("allreduce_variants"))
except Exception as e:
    print("$impl failed", e)
PY
done
