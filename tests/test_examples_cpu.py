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
    "moe_mnist": ["examples/moe/mnist_main.py", "--cpu", "--epochs", "2", "--steps-per-epoch", "3", "--num-local-experts", "2", "--set-deterministic", "--save-model",
                  "--log-interval", "1"],
    "moe_mnist_dense_async": ["examples/moe/mnist_main.py", "--cpu", "--epochs", "1", "--steps-per-epoch", "3", "--algorithm", "async", "--async-sync-interval", "5"],
    "primitives": ["examples/communication_primitives/main.py"],
    "elastic": ["examples/elastic_training/main.py", "--cpu", "--epochs", "1", "--steps-per-epoch", "3", "--checkpoint_path", "{tmp}/ckpt.pt", "--log-interval", "1",
                "--algorithm", "bytegrad"],
    "imagenet": ["examples/imagenet/main.py", "--cpu", "--synthetic", "--epochs", "1", "--steps-per-epoch", "2", "--batch-size", "2", "--num-classes", "10",
                 "--image-size", "32", "--print-freq", "1"],
    "synthetic_benchmark": ["examples/benchmark/synthetic_benchmark.py", "--cpu", "--model", "mnist", "--deterministic", "--num-warmup-batches", "1",
                            "--num-batches-per-iter", "2", "--num-iters", "2", "--batch-size", "4", "--algorithm", "low_precision_decentralized"],
    "squad": ["examples/squad/main.py", "--cpu", "--tiny", "--epochs", "1", "--num-synthetic", "24", "--batch-size", "4", "--max-seq-length", "64", "--doc_stride", "32",
              "--print-freq", "2", "--algorithm", "decentralized", "--output_dir", "{tmp}/out", "--do_lower_case", "--save_steps", "2", "--eval_all_checkpoints",
              "--gradient_accumulation_steps", "2", "--version_2_with_negative", "--evaluate_during_training"],
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_example_runs_under_the_static_launcher(name, tmp_path):
    from tests.mp_utils import free_port, run_in_session

    argv = [a.replace("{tmp}", str(tmp_path)) for a in CASES[name]]
    cmd = [sys.executable, "-m", "bagua_b200.distributed.launch", "--nproc_per_node=2", f"--master_port={free_port()}", os.path.join(REPO, argv[0]), *argv[1:]]
    r = run_in_session(cmd, 240, env=ENV, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]


def test_elastic_example_resumes_from_its_checkpoint(tmp_path):
    """examples/elastic_training/main.py: a second gang (here: a second launch, with a different world size) finds the first one's
    checkpoint, resumes with the following epoch and finishes; a third launch has nothing left to train."""
    from tests.mp_utils import free_port, run_in_session

    def launch(nproc, epochs):
        cmd = [sys.executable, "-m", "bagua_b200.distributed.launch", f"--nproc_per_node={nproc}", f"--master_port={free_port()}",
               os.path.join(REPO, "examples/elastic_training/main.py"), "--cpu", "--epochs", str(epochs), "--steps-per-epoch", "2", "--batch-size", "8",
               "--test-batch-size", "50", "--checkpoint_path", str(tmp_path / "ckpt.pt"), "--log-interval", "1"]
        r = run_in_session(cmd, 240, env=ENV, cwd=str(tmp_path))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        return r.stdout

    first = launch(2, 1)
    assert "resumed from" not in first and "Train Epoch: 1" in first
    second = launch(1, 2)
    assert "next epoch 2" in second and "Train Epoch: 2" in second and "Train Epoch: 1 " not in second
    third = launch(2, 2)
    assert "next epoch 3" in third and "nothing left to train" in third


def test_moe_example_final_loss_is_unchanged_by_a_checkpoint_round_trip(tmp_path):
    """The reference's CI runs examples/moe/mnist_main.py with and without --save-model under --set-deterministic and requires the
    same final loss (.buildkite/scripts/benchmark_master.sh:137-151): saving and re-loading model, experts and optimizer after
    every epoch must not change training."""
    import re

    from tests.mp_utils import free_port, run_in_session

    def final_loss(extra):
        cmd = [sys.executable, "-m", "bagua_b200.distributed.launch", "--nproc_per_node=2", f"--master_port={free_port()}",
               os.path.join(REPO, "examples/moe/mnist_main.py"), "--cpu", "--epochs", "2", "--steps-per-epoch", "3", "--num-local-experts", "2", "--set-deterministic",
               "--log-interval", "1", *extra]
        r = run_in_session(cmd, 240, env=ENV, cwd=str(tmp_path))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        return re.findall(r"Final Loss: ([0-9.]+)", r.stdout)[-1]

    assert final_loss([]) == final_loss(["--save-model", "--save-dir", str(tmp_path / "ckpt")])


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


_BARE = {k: v for k, v in ENV.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
_CONTRACT_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                  "config", "clocks", "e2e", "gpu_launches")


@pytest.mark.parametrize("ranks", [1, 2])
def test_bench_contract_on_the_host(ranks, tmp_path):
    """bench.py end to end in its host self-test mode, started exactly like the driver starts it: a bare ``python bench.py``
    (no launcher, no RANK/WORLD_SIZE in the environment) for one rank, ``torch.distributed.run`` for two.  One JSON line on
    stdout with every key of the contract; the e2e arm ran through DevicePrefetcher/LossReader."""
    import json

    from tests.mp_utils import free_port, run_in_session

    tail = ["--selftest-cpu", "--image-size", "32", "--batch-size", "2", "--steps", "2", "--warmup", "3", "--gpus", str(ranks)]
    if ranks == 1:
        cmd = [sys.executable, os.path.join(REPO, "bench.py"), *tail]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
               "--master-port", str(free_port()), os.path.join(REPO, "bench.py"), *tail]
    r = run_in_session(cmd, 600, env=dict(_BARE, OMP_NUM_THREADS="4"), cwd=str(tmp_path))  # torchrun would pin 1 thread per rank
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # only the result may reach stdout
    out = json.loads(lines[0])
    assert all(k in out for k in _CONTRACT_KEYS), sorted(out)
    assert out["n_gpus"] == ranks and out["steps"] == 2 and out["value"] > 0 and out["selftest"]
    assert out["e2e"]["value"] > 0 and out["e2e"]["h2d_bytes_per_step"] == 2 * 3 * 32 * 32 * 4 + 2 * 8 and out["e2e"]["d2h_bytes_per_step"] == 4
    assert out["config"]["global_batch"] == 2 * ranks and out["config"]["parallelism"] == f"dp{ranks}"
    # the post-measurement timeline probe (six untimed steps) reports how much communication stayed exposed after backward
    for block in (out, out["bert_large_bytegrad"]):
        tl = block["comm_timeline"]
        assert "error" not in tl and tl["steps"] >= 3 and tl["exposed_ms_median"] >= 0 and tl["buckets_per_step"] >= 1 and tl["bucket_finishing_last"] in tl["device_ms_per_bucket_median"]


def test_bench_reference_arm_reports_itself_unavailable(tmp_path):
    import json

    from tests.mp_utils import run_in_session

    r = run_in_session([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "3"], 120, env=_BARE,
                       cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["impl"] == "reference" and ("unavailable" in out or "value" in out)


def test_smoke_steps_without_a_launcher_environment(tmp_path):
    """``__graft_entry__.smoke()``'s training steps on the host, from a process with no RANK/WORLD_SIZE/MASTER_* (how the driver calls
    it): the default process group must come up as rank 0 of 1."""
    from tests.mp_utils import run_in_session

    code = ("import sys; sys.path.insert(0, %r)\n"
            "import torch, __graft_entry__ as g\n"
            "loss, opt = g._smoke_steps(torch.device('cpu'), torch.float32)\n"
            "import bagua_b200 as b\n"
            "assert b.get_rank() == 0 and b.get_world_size() == 1 and torch.isfinite(loss.detach()).item()\n"
            "print('SMOKE_HOST_OK')\n") % REPO
    r = run_in_session([sys.executable, "-c", code], 300, env=_BARE, cwd=str(tmp_path))
    assert r.returncode == 0 and "SMOKE_HOST_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_bench_verify_accepts_the_right_update_and_rejects_a_wrong_one():
    """bench.py's pre-timing check (fused bucket update vs an fp32-allreduce + hand-written SGD twin) on host doubles: an optimizer that
    applies exactly the rule passes with zero violations; bf16 rounding of the gradient before the update stays inside the tolerance;
    a wrong learning rate is refused."""
    import importlib.util
    import types

    import torch

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    c = types.SimpleNamespace(torch=torch, dist=torch.distributed, world=1, cpu=True, dtype=torch.bfloat16)

    def build():
        torch.manual_seed(0)
        return torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.ReLU(), torch.nn.Linear(64, 8)).to(torch.bfloat16)

    def loss_fn(m, x, y):
        return torch.nn.functional.cross_entropy(m(x).float(), y)

    class MasterSGD:
        """bf16 model, fp32 master weights; optionally rounds the gradient to bf16 first (what an in-switch bf16 sum does)."""

        def __init__(self, model, lr, round_grad=False):
            self.params = list(model.parameters())
            self.masters = [p.detach().float().clone() for p in self.params]
            self.lr, self.round_grad = lr, round_grad

        def zero_grad(self):
            for p in self.params:
                p.grad = None

        def step(self):
            with torch.no_grad():
                for p, m in zip(self.params, self.masters):
                    g = p.grad.float()
                    if self.round_grad:
                        g = (g * 8).to(torch.bfloat16).float() / 8
                    m.add_(g, alpha=-self.lr)
                    p.copy_(m)

    x, y = torch.randn(16, 32).to(torch.bfloat16), torch.randint(0, 8, (16,))
    for round_grad in (False, True):
        model = build()
        out = bench.verify_fused_update(c, build, model, MasterSGD(model, 0.5, round_grad), (x, y), loss_fn, 0.5)
        assert out["ok"] and out["fraction_outside_tolerance"] <= 1e-3, out
        if not round_grad:
            assert out["fraction_outside_tolerance"] == 0.0 and out["rel_l2_of_weight_diff_vs_update"] == 0.0
    model = build()
    with pytest.raises(SystemExit):
        bench.verify_fused_update(c, build, model, MasterSGD(model, 0.25), (x, y), loss_fn, 0.5)


def test_clock_sampler_regions_nearest_samples_and_one_shot_fallback(tmp_path, monkeypatch):
    """bench.py's nvidia-smi sampler against a stand-in ``nvidia-smi``: samples are selected by their own timestamps, a region shorter than
    the sampling period takes the nearest samples, throttle reasons are collected, and a looping sampler that delivers nothing (an
    nvidia-smi that rejects the interval) falls back to one query after the region."""
    import datetime
    import importlib.util
    import time

    spec = importlib.util.spec_from_file_location("bench_for_sampler_test", os.path.join(REPO, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    fake = tmp_path / "bin"
    fake.mkdir()
    smi = fake / "nvidia-smi"
    # looping mode (-lms): print nothing and exit (as if the interval were rejected); one-shot mode: one line stamped "now"
    smi.write_text("#!/bin/bash\nfor a in \"$@\"; do if [ \"$a\" = \"-lms\" ]; then exit 3; fi; done\n"
                   "echo \"$(date '+%Y/%m/%d %H:%M:%S.%3N'), 1965, 1965, 612.5, Not Active, Not Active, Not Active, Active\"\n")
    smi.chmod(0o755)
    monkeypatch.setenv("PATH", str(fake) + os.pathsep + os.environ["PATH"])
    s = bench.ClockSampler(0).start()
    time.sleep(0.3)
    t0 = time.time()
    rec = s.stop(t0 - 0.01, t0)
    assert rec["sm_mhz"] == 1965 and rec["sm_max_mhz"] == 1965 and rec["reasons"] == ["sw_power_cap"] and "one query" in rec["note"], rec

    # timestamp selection and the upper-half rule on injected samples
    s2 = bench.ClockSampler(0)
    s2.proc = object()
    base = time.time()

    def line(dt, mhz, thermal="Not Active"):
        ts = datetime.datetime.fromtimestamp(base + dt).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return f"{ts}, {mhz}, 1965, 500.0, Not Active, {thermal}, Not Active, Not Active\n"

    s2._buf = [line(-5.0, 300), line(0.1, 1900), line(0.2, 1965), line(0.3, 1965), line(0.4, 1950), line(9.0, 210, "Active")]
    inside = s2.summary(base, base + 0.5)
    assert inside["samples"] == 4 and inside["sm_mhz"] == 1965 and inside["reasons"] == [] and "note" not in inside
    short = s2.summary(base + 2.0, base + 2.01)
    assert short["samples"] == 2 and "nearest samples" in short["note"]
    hot = s2.summary(base + 8.9, base + 9.1)
    assert hot["reasons"] == ["hw_thermal_slowdown"] and hot["sm_mhz"] == 210
