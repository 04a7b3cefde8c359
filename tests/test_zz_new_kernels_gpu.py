"""GPU tests of kernels added after the last hardware session (they sort last on purpose: the round-end run uses ``-x``, and
a surprise here must not hide the results of the validated suites).  Same oracle rule as test_kernels_gpu.py: compare
against a plain fp32 PyTorch implementation of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def test_fused_adam_mixed_precision_multi_tensor(dev):
    """bf16 parameters that are not bucket-flattened: fp32 moments + fp32 master weights in the multi-tensor kernel."""
    from bagua_b200.ops.optim import FusedAdam

    torch.manual_seed(9)
    ps = [torch.nn.Parameter((torch.randn(s, device=dev) * 0.1).to(torch.bfloat16)) for s in [(65, 33), (1000,), (7, 9, 11)]]
    rs = [torch.nn.Parameter(p.detach().float().clone()) for p in ps]
    opt, ropt = FusedAdam(ps, lr=1e-3, adamw=True, weight_decay=0.01), torch.optim.AdamW(rs, lr=1e-3, weight_decay=0.01)
    for _ in range(20):
        for p, r in zip(ps, rs):
            g = torch.randn_like(r)
            p.grad, r.grad = g.to(torch.bfloat16), g.to(torch.bfloat16).float()
        opt.step()
        ropt.step()
    for p, r in zip(ps, rs):
        assert opt.state[p]["exp_avg"].dtype == torch.float32 and "master" in opt.state[p]
        torch.testing.assert_close(opt.state[p]["master"], r.data, rtol=1e-4, atol=1e-5)   # fp32 trajectory is preserved
        torch.testing.assert_close(p.data.float(), r.data, rtol=0, atol=8e-3)              # parameter = rounded master
