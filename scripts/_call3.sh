mkdir -p gpurun_out/c3
timeout 300 python -m pytest tests/test_self_peer_gpu.py -q --timeout 300 -p no:cacheprovider -k "moe" > gpurun_out/c3/pytest_moe.log 2>&1
tail -3 gpurun_out/c3/pytest_moe.log
run() { name=$1; shift; timeout 300 env "$@" python bench.py --gpus 1 --steps 30 --warmup 5 --workloads vgg16 $EXTRA > gpurun_out/c3/b_$name.json 2> gpurun_out/c3/b_$name.err; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/c3/b_$name.json")); print("$name", round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],3), d["gpu_launches"], d["config"]["allreduce_variants"])
except Exception as e: print("$name failed", e)
PY
}
EXTRA="" run default A=1
EXTRA="" run lag4 BAGUA_BENCH_LOSS_LAG=4
EXTRA="" run inline BAGUA_INLINE_COMM=1
EXTRA="--no-self-peer" run plain A=1
EXTRA="" run blocks8 BAGUA_FUSED_BLOCKS=8
EXTRA="" run blocks64 BAGUA_FUSED_BLOCKS=64
EXTRA="" run nativehooks BAGUA_NATIVE_HOOKS=1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:allreduce_|all_gather_k|reduce_scatter_k|peer_average|bytegrad|lpdec|async_average|flat_sgd|flat_adam|minmax_uint8|bias_relu|moe_|grouped_gemm' -o gpurun_out/c3/ncu_zoo python scripts/kernel_zoo.py > gpurun_out/c3/ncu_zoo.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c3/ncu_zoo.log; ls -la gpurun_out/c3 | grep ncu
