mkdir -p gpurun_out/c4
for cfg in gpt2_moe resnet50_decentralized resnet50_async; do
 for arm in peer nccl; do
  timeout 300 python benchmarks/config_bench.py --config $cfg --arm $arm --steps 10 --warmup 3 > gpurun_out/c4/${cfg}_${arm}.json 2> gpurun_out/c4/${cfg}_${arm}.err
  echo "$cfg $arm rc=$?"; tail -c 600 gpurun_out/c4/${cfg}_${arm}.json; grep -n "Error" gpurun_out/c4/${cfg}_${arm}.err | tail -3
 done
done
BAGUA_MOE_FUSED_COMBINE=1 timeout 300 python benchmarks/config_bench.py --config gpt2_moe --arm peer --steps 10 --warmup 3 > gpurun_out/c4/gpt2_moe_peer_fusedcombine.json 2> gpurun_out/c4/gpt2_moe_peer_fusedcombine.err; echo "fused combine rc=$?"; tail -c 400 gpurun_out/c4/gpt2_moe_peer_fusedcombine.json; grep -n "Error" gpurun_out/c4/gpt2_moe_peer_fusedcombine.err | tail -3
timeout 600 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider -x > gpurun_out/c4/pytest_gpu.log 2>&1; tail -5 gpurun_out/c4/pytest_gpu.log
