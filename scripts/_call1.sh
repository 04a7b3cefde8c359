mkdir -p gpurun_out/c1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1/smi.txt 2>&1
timeout 120 python scripts/probe_self_peer.py > gpurun_out/c1/probe.json 2> gpurun_out/c1/probe.err
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/c1/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c1/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/c1/smoke.log
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c1/bench_n1.json 2> gpurun_out/c1/bench_n1.err
echo "bench rc=$?" >> gpurun_out/c1/bench_n1.err
tail -5 gpurun_out/c1/pytest_gpu.log; cat gpurun_out/c1/probe.json; tail -2 gpurun_out/c1/smoke.log; cat gpurun_out/c1/bench_n1.json | head -c 1500
