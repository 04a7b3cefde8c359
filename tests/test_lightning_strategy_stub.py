"""BaguaStrategy against a stub of the Lightning strategy interface.

Lightning is not installed in this image (and the reference's integration test needs it: tests/pytorch_lightning/
test_bagua_strategy.py:30-107 fits a BoringModel with ``Trainer(strategy=BaguaStrategy(algorithm=...))``).  What the strategy adds on
top of Lightning's DDPStrategy is small and testable without it: process-group set-up from the cluster environment, wrapping the
LightningModule so that ``training_step`` runs under the engine's hooks, building the algorithm (incl. QAdam's optimizer check),
refusing torch-DDP wrapping, stopping the asynchronous averaging thread at teardown.  The stub below provides exactly the attributes
and call order of ``DDPStrategy`` that the strategy touches; the test then drives the strategy the way ``Trainer.fit`` does."""
import os
import subprocess
import sys
import textwrap

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

STUB = '''
import types
import torch


class _Env:
    main_address, main_port = "127.0.0.1", {port}
    def global_rank(self): return 0
    def world_size(self): return 1
    def local_rank(self): return 0
    def node_rank(self): return 0


class _Accelerator:
    def setup(self, trainer): self.was_set_up = True


class DDPStrategy:
    """The slice of lightning.pytorch.strategies.DDPStrategy that BaguaStrategy relies on."""
    def __init__(self, accelerator=None, parallel_devices=None, cluster_environment=None, checkpoint_io=None, precision_plugin=None):
        self.accelerator = accelerator or _Accelerator()
        self.cluster_environment = cluster_environment or _Env()
        self.root_device = torch.device("cpu")
        self.num_processes = 1
        self.lightning_module = None
        self.optimizers = []
        self.model = None
        self.torn_down = False
    def model_to_device(self): self.lightning_module.to(self.root_device)
    def setup_optimizers(self, trainer): self.optimizers = [self.lightning_module.configure_optimizers()]
    def setup_precision_plugin(self): pass
    def configure_ddp(self): raise AssertionError("torch DDP wrapping must be overridden")
    def teardown(self): self.torn_down = True
'''

DRIVER = '''
import sys, types, torch
sys.path.insert(0, {repo!r}); sys.path.insert(0, {stub_dir!r})
import bagua_b200 as bagua
from bagua_b200.contrib.lightning import BaguaStrategy, lightning_available
from bagua_b200.parallel.algorithms.q_adam import QAdamOptimizer
assert lightning_available()


class Boring(torch.nn.Module):   # a LightningModule as far as the strategy is concerned
    def __init__(self, qadam=False):
        super().__init__()
        self.layer = torch.nn.Linear(32, 2)
        self.qadam = qadam
        self.steps_seen = 0
    def forward(self, x): return self.layer(x)
    def training_step(self, batch, batch_idx):
        self.steps_seen += 1
        return self(batch).pow(2).mean()
    def validation_step(self, batch, batch_idx): return self(batch).mean()
    def configure_optimizers(self):
        return QAdamOptimizer(self.parameters(), lr=1e-2, warmup_steps=2) if self.qadam else torch.optim.SGD(self.parameters(), lr=0.1)


def fit(algorithm, qadam=False, **kw):
    module = Boring(qadam)
    trainer = types.SimpleNamespace(state=types.SimpleNamespace(fn="TrainerFn.FITTING"), training=True, testing=False, sanity_checking=False,
                                    validating=False, predicting=False)
    module._trainer = trainer
    strategy = BaguaStrategy(algorithm=algorithm, **kw)
    strategy.lightning_module = module
    strategy.setup_distributed()
    assert bagua.is_initialized() and bagua.get_world_size() == 1
    strategy.setup(trainer)
    strategy.configure_ddp()                                   # must be a no-op, not torch DDP
    from bagua_b200.parallel.data_parallel.distributed import DistributedDataParallel_V1_9_0_Interface
    assert isinstance(strategy.model, DistributedDataParallel_V1_9_0_Interface), type(strategy.model)
    before = module.layer.weight.detach().clone()
    opt = strategy.optimizers[0]
    for i in range(4):
        opt.zero_grad()
        loss = strategy.model(torch.randn(8, 32), i)           # Lightning calls the wrapped model; it must land in training_step
        loss.backward()
        opt.step()
    assert module.steps_seen == 4 and not torch.equal(before, module.layer.weight)
    trainer.training, trainer.validating = False, True
    assert strategy.model(torch.randn(8, 32), 0).dim() == 0    # routed to validation_step
    strategy.teardown()
    assert strategy.torn_down
    return strategy


for name in ("gradient_allreduce", "bytegrad", "decentralized", "async"):
    s = fit(name, **({{"sync_interval_ms": 5}} if name == "async" else {{}}))
    print("OK", name, type(s.model.inner.bagua_algorithm).__name__)
fit("qadam", qadam=True)
print("OK qadam")
try:
    fit("qadam", qadam=False)
    raise SystemExit("qadam without a QAdamOptimizer must be refused")
except ValueError as e:
    print("OK refused:", e)
'''


def test_bagua_strategy_drives_the_engine_through_a_stub_of_lightning(tmp_path):
    from tests.mp_utils import free_port, run_in_session

    pkg = tmp_path / "pytorch_lightning"
    (pkg / "strategies").mkdir(parents=True)
    (pkg / "utilities").mkdir()
    (pkg / "__init__.py").write_text("")
    (pkg / "strategies" / "__init__.py").write_text(STUB.format(port=free_port()))
    (pkg / "utilities" / "__init__.py").write_text("")
    (pkg / "utilities" / "optimizer.py").write_text("def _optimizers_to_device(optimizers, device):\n    return optimizers\n")
    driver = tmp_path / "drive.py"
    driver.write_text(textwrap.dedent(DRIVER.format(repo=REPO, stub_dir=str(tmp_path))))
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
    env.update(CUDA_VISIBLE_DEVICES="", BAGUA_FORCE_CPU="1")
    r = run_in_session([sys.executable, str(driver)], 300, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for name in ("gradient_allreduce", "bytegrad", "decentralized", "async", "qadam", "refused"):
        assert f"OK {name}" in r.stdout, r.stdout
