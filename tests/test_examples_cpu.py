"""The shipped examples as acceptance workloads: each runs for a few steps on CPU/gloo with two
processes under the static launcher, exactly as a user would start it."""
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = dict(os.environ, PYTHONPATH=REPO + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="", BAGUA_FORCE_CPU="1")

CASES = {
    "mnist": ["examples/mnist/main.py", "--cpu", "--epochs", "1", "--steps-per-epoch", "5", "--algorithm", "bytegrad"],
    "mnist_fused": ["examples/mnist/main.py", "--cpu", "--epochs", "1", "--steps-per-epoch", "5", "--fuse-optimizer"],
    "moe_mnist": ["examples/moe/mnist_main.py", "--cpu", "--steps", "5"],
    "primitives": ["examples/communication_primitives/main.py"],
    "elastic": ["examples/elastic_training/main.py", "--cpu", "--steps", "5", "--ckpt", "{tmp}/ckpt.pt"],
    "imagenet": ["examples/imagenet/main.py", "--cpu", "--synthetic", "--epochs", "1", "--steps-per-epoch", "2", "--batch-size", "2", "--num-classes", "10",
                 "--image-size", "32", "--print-freq", "1"],
    "synthetic_benchmark": ["examples/benchmark/synthetic_benchmark.py", "--cpu", "--model", "mnist", "--deterministic", "--num-warmup-batches", "1",
                            "--num-batches-per-iter", "2", "--num-iters", "2", "--batch-size", "4", "--algorithm", "low_precision_decentralized"],
    "squad": ["examples/squad/main.py", "--cpu", "--tiny", "--epochs", "1", "--num-synthetic", "48", "--batch-size", "4", "--max-seq-length", "64",
              "--print-freq", "2", "--algorithm", "decentralized"],
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_example_runs_under_the_static_launcher(name, tmp_path):
    from tests.mp_utils import free_port, run_in_session

    argv = [a.replace("{tmp}", str(tmp_path)) for a in CASES[name]]
    cmd = [sys.executable, "-m", "bagua_b200.distributed.launch", "--nproc_per_node=2", f"--master_port={free_port()}", os.path.join(REPO, argv[0]), *argv[1:]]
    r = run_in_session(cmd, 240, env=ENV, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


@pytest.mark.parametrize("config", ["bert_bytegrad", "resnet50_decentralized", "resnet50_async", "gpt2_moe"])
def test_baseline_config_benchmarks_run_tiny_in_bf16(config, tmp_path):
    """benchmarks/config_bench.py for every BASELINE configuration: tiny models, two CPU ranks, bf16 parameters — the python
    paths of the GPU runs (dtype handling, algorithm wiring, MoE gate) without spending GPU time on a typo."""
    import json

    from tests.mp_utils import free_port, run_in_session

    cmd = [sys.executable, "-m", "bagua_b200.distributed.launch", "--nproc_per_node=2", f"--master_port={free_port()}",
           os.path.join(REPO, "benchmarks/config_bench.py"), "--config", config, "--tiny", "--cpu", "--force-bf16", "--steps", "2", "--warmup", "3", "--batch-size", "2"]
    r = run_in_session(cmd, 300, env=ENV, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["config"] == config and out["n_gpus"] == 2 and out["loss_finite"] and out["value"] > 0
