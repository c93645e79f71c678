#!/usr/bin/env python
"""Writes tests/golden/adam_*.npz from torch.optim.Adam itself (CPU, fp32), constructed the way the reference does
(scene/gaussian_model.py:197-206: six groups with their own learning rates, lr=0.0 default, eps=1e-15), stepped a few
times with seeded gradients and a per-step change of the first group's learning rate (update_learning_rate).

    python oracle/make_golden_adam.py
"""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GROUPS = [("xyz", (3,), 1.6e-4), ("f_dc", (1, 3), 2.5e-3), ("f_rest", (15, 3), 2.5e-3 / 20.0), ("opacity", (1,), 0.05),
          ("scaling", (3,), 5e-3), ("rotation", (4,), 1e-3)]


def make(P, steps, seed, path):
    g = torch.Generator().manual_seed(seed)
    params = [torch.nn.Parameter(torch.randn(P, *shape, generator=g)) for _, shape, _ in GROUPS]
    opt = torch.optim.Adam([{"params": [p], "lr": lr, "name": name} for p, (name, _, lr) in zip(params, GROUPS)],
                           lr=0.0, eps=1e-15)
    out = {"P": P, "steps": steps}
    for (name, _, _), p in zip(GROUPS, params):
        out[f"p0_{name}"] = p.detach().numpy().copy()
    for t in range(steps):
        for gi, ((name, shape, lr), p) in enumerate(zip(GROUPS, params)):
            scale = 10.0 ** float(torch.randint(-6, 1, (1,), generator=g))       # gradients of very different sizes
            grad = torch.randn(P, *shape, generator=g) * scale
            if t == 1:
                grad[: P // 3] = 0.0                                              # culled rows: zero gradient
            p.grad = grad
            out[f"g{t}_{name}"] = grad.numpy().copy()
        opt.param_groups[0]["lr"] = GROUPS[0][2] * (0.9 ** t)                    # update_learning_rate(iteration)
        out[f"lr{t}"] = np.array([grp["lr"] for grp in opt.param_groups], dtype=np.float64)
        opt.step()
    for (name, _, _), p in zip(GROUPS, params):
        st = opt.state[p]
        out[f"p_{name}"] = p.detach().numpy().copy()
        out[f"m_{name}"] = st["exp_avg"].numpy().copy()
        out[f"v_{name}"] = st["exp_avg_sq"].numpy().copy()
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    make(130, 4, 11, os.path.join(ROOT, "tests", "golden", "adam_small.npz"))
    make(61, 7, 12, os.path.join(ROOT, "tests", "golden", "adam_7steps.npz"))
