"""CPU: the Adam restatement (oracle/adam_oracle.py) against torch.optim.Adam itself and against the committed
goldens (written by oracle/make_golden_adam.py from torch.optim.Adam with the reference's constructor arguments)."""
import glob
import os

import numpy as np
import pytest
import torch

from parity_common import ROOT
from oracle import adam_oracle as AO

GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "adam_*.npz")))
NAMES = ["xyz", "f_dc", "f_rest", "opacity", "scaling", "rotation"]


@pytest.mark.parametrize("path", GOLD, ids=os.path.basename)
def test_oracle_against_torch_adam_golden(path):
    z = np.load(path)
    steps = int(z["steps"])
    for gi, name in enumerate(NAMES):
        grads = [z[f"g{t}_{name}"] for t in range(steps)]
        lrs = [float(z[f"lr{t}"][gi]) for t in range(steps)]
        p, m, v = AO.adam_run(z[f"p0_{name}"], grads, lrs, eps=1e-15)
        # the golden is fp32 arithmetic, the oracle float64: a few fp32 ulps per step
        assert np.abs(p - z[f"p_{name}"]).max() <= 2e-6 * max(1.0, np.abs(z[f"p_{name}"]).max())
        assert np.abs(m - z[f"m_{name}"]).max() <= 1e-6 * max(1e-30, np.abs(z[f"m_{name}"]).max())
        assert np.abs(v - z[f"v_{name}"]).max() <= 1e-6 * max(1e-30, np.abs(z[f"v_{name}"]).max())


def test_oracle_against_live_torch_adam_float64():
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(300, 7, generator=g, dtype=torch.float64)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([p], lr=3e-3, eps=1e-15)
    grads = []
    for t in range(5):
        gr = torch.randn(300, 7, generator=g, dtype=torch.float64) * (10.0 ** (t - 4))
        grads.append(gr.numpy().copy())
        p.grad = gr
        opt.step()
    want_p, want_m, want_v = AO.adam_run(p0.numpy(), grads, [3e-3] * 5, eps=1e-15)
    st = opt.state[p]
    assert np.abs(want_p - p.detach().numpy()).max() <= 1e-13
    assert np.abs(want_m - st["exp_avg"].numpy()).max() <= 1e-15
    assert np.abs(want_v - st["exp_avg_sq"].numpy()).max() <= 1e-15


def test_bags_adam_host_logic():
    """Constructor checks, group handling and the refusal to run on CPU tensors (no GPU needed)."""
    import bags_optim
    w = torch.nn.Parameter(torch.zeros(4, 3))
    with pytest.raises(ValueError):
        bags_optim.BagsAdam([w], lr=1e-3, weight_decay=0.1)
    with pytest.raises(ValueError):
        bags_optim.BagsAdam([w], lr=1e-3, amsgrad=True)
    with pytest.raises(ValueError):
        bags_optim.BagsAdam([w], lr=-1.0)
    opt = bags_optim.BagsAdam([{"params": [w], "lr": 0.01, "name": "xyz"}], lr=0.0, eps=1e-15)
    assert opt.param_groups[0]["name"] == "xyz" and opt.param_groups[0]["lr"] == 0.01
    assert opt.param_groups[0]["eps"] == 1e-15 and opt.param_groups[0]["betas"] == (0.9, 0.999)
    opt.step()                                   # no gradients: nothing to do, no error
    w.grad = torch.ones_like(w)
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step()
    # state_dict layout interoperates with torch.optim.Adam
    ref = torch.optim.Adam([{"params": [w], "lr": 0.01, "name": "xyz"}], lr=0.0, eps=1e-15)
    assert set(opt.state_dict()["param_groups"][0]) >= {"lr", "betas", "eps", "name", "params"}
    ref.load_state_dict(opt.state_dict())
