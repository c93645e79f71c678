"""Fused Adam: drop-in for the ``torch.optim.Adam`` the reference builds in ``GaussianModel.training_setup``
(scene/gaussian_model.py:197-206, stepped at train.py:420), SURVEY.md section 8f row N4.

    self.optimizer = BagsAdam(l, lr=0.0, eps=1e-15)          # instead of torch.optim.Adam(l, lr=0.0, eps=1e-15)

Same constructor arguments (no weight decay / amsgrad / maximize: the reference uses none of them), same parameter
groups (per-group ``lr``, ``name`` and any other key are kept), same ``state`` layout (``step`` as a CPU fp32 scalar
tensor, ``exp_avg``, ``exp_avg_sq``), so the reference's densification code -- which reads, concatenates, masks and
replaces ``optimizer.state[param]["exp_avg"/"exp_avg_sq"]`` (scene/gaussian_model.py:316-400) -- and
``update_learning_rate`` keep working unchanged, and ``state_dict`` / ``load_state_dict`` interoperate with
``torch.optim.Adam``.

``step()`` runs the whole update of up to eight tensors as ONE kernel of ``libbags_b200.so``
(``csrc/adam.cu``, C ABI ``bags_adam_step``): 28 bytes of traffic per element instead of the ~100 of PyTorch's nine
``_foreach_*`` passes.  The arithmetic follows PyTorch's foreach implementation operation by operation.  There is no
CPU path: CPU parameters raise."""
import ctypes

import torch

import bags_capi as capi


class BagsAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, *,
                 maximize=False):
        if weight_decay != 0 or amsgrad or maximize:
            raise ValueError("BagsAdam implements the configuration the reference uses: weight_decay=0, "
                             "amsgrad=False, maximize=False")
        if isinstance(lr, torch.Tensor):
            raise ValueError("BagsAdam: lr must be a Python number")
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = capi.lib()
        # tensors are batched per (device, betas, eps): one bags_adam_step call each
        batches = {}
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            lr = group["lr"]
            if isinstance(lr, torch.Tensor):
                lr = float(lr)
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda:
                    raise RuntimeError("bags_b200: BagsAdam needs CUDA parameters (there is no CPU path)")
                if p.grad.is_sparse:
                    raise RuntimeError("BagsAdam does not support sparse gradients")
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("BagsAdam: parameters must be contiguous fp32 tensors")
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                step = float(state["step"])
                grad = p.grad
                if grad.dtype != torch.float32 or not grad.is_contiguous():
                    grad = grad.to(torch.float32).contiguous()
                m, v = state["exp_avg"], state["exp_avg_sq"]
                if not (m.is_contiguous() and v.is_contiguous() and m.dtype == torch.float32
                        and v.dtype == torch.float32 and m.device == p.device and v.device == p.device):
                    raise RuntimeError("BagsAdam: exp_avg / exp_avg_sq must be contiguous fp32 tensors on the "
                                       "parameter's device")
                bias_correction1 = 1 - beta1 ** step
                bias_correction2 = 1 - beta2 ** step
                t = capi.AdamTensor(p.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(),
                                    (lr / bias_correction1) * -1, bias_correction2 ** 0.5)
                batches.setdefault((p.device, beta1, beta2, group["eps"]), []).append((t, grad))
        for (dev, beta1, beta2, eps), items in batches.items():
            arr = (capi.AdamTensor * len(items))(*[t for t, _ in items])
            with torch.cuda.device(dev):
                stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                rc = L.bags_adam_step(len(items), ctypes.cast(arr, ctypes.c_void_p), float(beta1), float(beta2),
                                      float(eps), stream)
            capi.check(rc, "bags_adam_step")
        return loss
