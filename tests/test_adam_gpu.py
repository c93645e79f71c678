"""GPU parity of the fused Adam step (csrc/adam.cu through bags_optim.BagsAdam) against
torch.optim.Adam: the committed goldens (torch CPU, the reference's constructor arguments), the float64 oracle, and
torch's own CUDA implementation on the same device.  Tolerance: 1e-6 relative (fp32 arithmetic on both sides; the
kernel follows torch's foreach operation order, so most elements are bit-identical)."""
import glob
import os

import numpy as np
import pytest
import torch

from parity_common import ROOT
from oracle import adam_oracle as AO

pytestmark = pytest.mark.gpu
GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "adam_*.npz")))
NAMES = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"]


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _close(a, b, rel=1e-6):
    a, b = a.double().cpu(), b.double().cpu()
    scale = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) <= rel * scale


@pytest.mark.parametrize("path", GOLD, ids=os.path.basename)
def test_adam_against_torch_golden(path):
    import bags_optim
    import bags_capi
    dev = _dev()
    z = np.load(path)
    steps = int(z["steps"])
    params = [torch.nn.Parameter(torch.from_numpy(z[f"p0_{n}"]).to(dev)) for n in NAMES]
    opt = bags_optim.BagsAdam([{"params": [p], "lr": float(z["lr0"][i]), "name": n}
                               for i, (p, n) in enumerate(zip(params, NAMES))], lr=0.0, eps=1e-15)
    n0 = bags_capi.lib().bags_launch_count()
    for t in range(steps):
        for p, n in zip(params, NAMES):
            p.grad = torch.from_numpy(z[f"g{t}_{n}"]).to(dev)
        for i, grp in enumerate(opt.param_groups):
            grp["lr"] = float(z[f"lr{t}"][i])
        opt.step()
    assert bags_capi.lib().bags_launch_count() == n0 + steps          # ONE launch per step for the six tensors
    for gi, (p, n) in enumerate(zip(params, NAMES)):
        st = opt.state[p]
        assert float(st["step"]) == steps
        assert _close(p.detach(), torch.from_numpy(z[f"p_{n}"]))
        assert _close(st["exp_avg"], torch.from_numpy(z[f"m_{n}"]))
        assert _close(st["exp_avg_sq"], torch.from_numpy(z[f"v_{n}"]))
        # and the float64 oracle
        grads = [z[f"g{t}_{n}"] for t in range(steps)]
        lrs = [float(z[f"lr{t}"][gi]) for t in range(steps)]
        op, om, ov = AO.adam_run(z[f"p0_{n}"], grads, lrs, eps=1e-15)
        assert _close(p.detach(), torch.from_numpy(op), rel=2e-6)


def test_adam_against_live_torch_cuda_flat_buffer():
    """The view-parallel layout: parameters and gradients are views into two flat buffers (P not a multiple of 4, so
    most views are not 16-byte aligned), eleven tensors (two launches), an empty tensor, one without gradient."""
    import bags_optim
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    P = 100_003
    shapes = [(P, 3), (P, 16, 3), (P, 1), (P, 3), (P, 4), (7,), (0, 3), (5, 5), (1,), (1023,), (1025,)]
    lrs = [1.6e-4, 2.5e-3, 0.05, 5e-3, 1e-3, 0.1, 0.1, 1e-2, 1.0, 1e-3, 1e-3]
    total = sum(int(np.prod(s)) for s in shapes)
    flat0 = torch.randn(total, generator=g)

    def build(optim_cls):
        flat = flat0.clone().to(dev)
        flat_grad = torch.zeros_like(flat)
        ps, off = [], 0
        for s in shapes:
            n = int(np.prod(s))
            p = flat[off:off + n].view(*s).detach().requires_grad_(True)
            p.grad = flat_grad[off:off + n].view(*s)
            ps.append(p)
            off += n
        extra = torch.nn.Parameter(torch.ones(9, device=dev))          # never receives a gradient
        opt = optim_cls([{"params": [p], "lr": lr} for p, lr in zip(ps, lrs)] + [{"params": [extra], "lr": 1.0}],
                        lr=0.0, eps=1e-15)
        return flat, flat_grad, ps, extra, opt

    fa, ga, pa, ea, oa = build(bags_optim.BagsAdam)
    fb, gb, pb, eb, ob = build(torch.optim.Adam)
    for t in range(4):
        grad = torch.randn(total, generator=g) * (10.0 ** (t - 3))
        grad[::7] = 0.0
        ga.copy_(grad.to(dev))
        gb.copy_(grad.to(dev))
        oa.param_groups[0]["lr"] = ob.param_groups[0]["lr"] = 1.6e-4 * (0.5 ** t)
        oa.step()
        ob.step()
    torch.cuda.synchronize()
    assert _close(fa, fb)
    same = float((fa == fb).float().mean())
    assert same > 0.99, same                                   # same operation order: (nearly) every element identical
    for x, y in zip(pa, pb):
        if x.numel() == 0:
            continue
        assert _close(oa.state[x]["exp_avg"], ob.state[y]["exp_avg"])
        assert _close(oa.state[x]["exp_avg_sq"], ob.state[y]["exp_avg_sq"])
    assert torch.equal(ea, eb) and len(oa.state[ea]) == 0


def test_adam_state_is_compatible_with_reference_densification():
    """scene/gaussian_model.py:316-400 replaces / concatenates / masks `optimizer.state[p]["exp_avg"]`: the same
    manipulations on BagsAdam's state keep working and the next step uses the new tensors."""
    import bags_optim
    dev = _dev()
    p = torch.nn.Parameter(torch.randn(10, 3, device=dev))
    opt = bags_optim.BagsAdam([{"params": [p], "lr": 0.01, "name": "xyz"}], lr=0.0, eps=1e-15)
    p.grad = torch.randn_like(p)
    opt.step()
    group = opt.param_groups[0]
    stored = opt.state.get(group["params"][0], None)
    ext = torch.randn(4, 3, device=dev)
    stored["exp_avg"] = torch.cat((stored["exp_avg"], torch.zeros_like(ext)), dim=0)          # cat_tensors_to_optimizer
    stored["exp_avg_sq"] = torch.cat((stored["exp_avg_sq"], torch.zeros_like(ext)), dim=0)
    del opt.state[group["params"][0]]
    group["params"][0] = torch.nn.Parameter(torch.cat((group["params"][0], ext), dim=0).requires_grad_(True))
    opt.state[group["params"][0]] = stored
    q = group["params"][0]
    q.grad = torch.randn_like(q)
    before = q.detach().clone()
    opt.step()
    assert float(opt.state[q]["step"]) == 2
    assert float((q.detach() - before).abs().max()) > 0
    with pytest.raises(RuntimeError):
        bad = torch.nn.Parameter(torch.zeros(3))
        o2 = bags_optim.BagsAdam([bad], lr=0.1)
        bad.grad = torch.ones(3)
        o2.step()
