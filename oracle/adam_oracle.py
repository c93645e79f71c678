"""TEST INFRASTRUCTURE ONLY -- host restatement of the optimizer step on the path.

Nothing in the product (`bundle-adjusting-gaussian-splatting_b200/`) imports this file; the product path is the CUDA
kernel of csrc/adam.cu and has no CPU fallback.

The reference does not implement Adam itself: `GaussianModel.training_setup` (scene/gaussian_model.py:197-206)
instantiates `torch.optim.Adam(l, lr=0.0, eps=1e-15)` -- a third-party dependency (PyTorch; the reference's
environment.yml pins pytorch 1.12.1, this image has 2.11.0) -- over six parameter groups, and train.py:420 calls
`.step()`.  This file restates the published algorithm (Kingma & Ba 2015, Algorithm 1, in the form PyTorch documents
for torch.optim.Adam with weight_decay=0, amsgrad=False, maximize=False):

    m_t = beta1 m_{t-1} + (1 - beta1) g_t
    v_t = beta2 v_{t-1} + (1 - beta2) g_t^2
    theta_t = theta_{t-1} - [lr / (1 - beta1^t)] * m_t / (sqrt(v_t) / sqrt(1 - beta2^t) + eps)

in float64 numpy.  Parity pinning: tests/test_adam_oracle_golden.py checks it against `torch.optim.Adam` itself run
on the CPU in this container (fp32 and fp64) and against tests/golden/adam_*.npz, written by oracle/make_golden_adam.py
from `torch.optim.Adam` with the reference's own constructor arguments and group learning rates.  Status: pinned.
"""
import numpy as np


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """One step for one tensor; `step` is the 1-based count of this update.  Returns (param, exp_avg, exp_avg_sq) as
    float64 arrays."""
    p = np.asarray(param, dtype=np.float64)
    g = np.asarray(grad, dtype=np.float64)
    m = np.asarray(exp_avg, dtype=np.float64)
    v = np.asarray(exp_avg_sq, dtype=np.float64)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    bias_correction1 = 1.0 - beta1 ** step
    bias_correction2 = 1.0 - beta2 ** step
    denom = np.sqrt(v) / np.sqrt(bias_correction2) + eps
    p = p - (lr / bias_correction1) * (m / denom)
    return p, m, v


def adam_run(param0, grads, lrs, beta1=0.9, beta2=0.999, eps=1e-8):
    """len(grads) consecutive steps from zero state; lrs[t] is the learning rate of step t (the reference changes the
    xyz group's lr every iteration, scene/gaussian_model.py:212-218)."""
    p = np.asarray(param0, dtype=np.float64)
    m = np.zeros_like(p)
    v = np.zeros_like(p)
    for t, (g, lr) in enumerate(zip(grads, lrs), start=1):
        p, m, v = adam_step(p, g, m, v, t, lr, beta1, beta2, eps)
    return p, m, v
