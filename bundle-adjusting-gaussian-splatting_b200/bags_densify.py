"""Fused densification: drop-in for ``GaussianModel.densify_and_prune`` of the reference
(scene/gaussian_model.py:433-447 -> densify_and_clone :416-431, densify_and_split :395-414, prune_points :334-348,
densification_postfix :375-393, cat_tensors_to_optimizer / _prune_optimizer :315-373), SURVEY.md section 8f row N4.

    import bags_densify
    bags_densify.densify_and_prune(gaussians, max_grad, min_opacity, extent, max_screen_size)

does to ``gaussians`` (a reference ``GaussianModel`` with its Adam optimizer) what the reference method does: the same
rows survive, in the same order, with the same values and optimizer state.  The reference performs three rounds of
``torch.cat`` / boolean indexing over every parameter tensor and both Adam moments (about seventy full-size passes and a
dozen host synchronisations); here the three selection masks are evaluated with the reference's own expressions
(element-wise work on P rows), exclusive-scanned, and ONE kernel (``bags_densify_gather``, csrc/densify.cu) moves every
surviving row of every tensor exactly once.

Random numbers: ``densify_and_split`` draws ``torch.normal(mean=0, std=scaling[selected].repeat(N, 1))``; the same call
is made here, first, so a seeded generator yields the same samples as in the reference.
There is no CPU path: CPU tensors raise.
"""
import ctypes

import torch
from torch import nn

import bags_capi as capi

# optimizer group name -> GaussianModel attribute (scene/gaussian_model.py:197-204, 338-343)
GROUP_ATTR = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity",
              "scaling": "_scaling", "rotation": "_rotation"}
ROLE = {"xyz": 1, "scaling": 2}


def selection_masks(g, max_grad, min_opacity, extent, max_screen_size, N=2):
    """The reference's selection logic on the P rows the model has BEFORE densification, as four boolean vectors:
    keep (row survives as itself), clone (its copy survives), child (its N split children survive), split (it is a split
    parent: removed, and a sample is drawn for it whether or not its children survive).  Every expression is the
    reference's own (line numbers of scene/gaussian_model.py)."""
    grads = g.xyz_gradient_accum / g.denom                                                  # :434
    grads[grads.isnan()] = 0.0                                                              # :435
    scale_max = torch.max(g.get_scaling, dim=1).values
    small = scale_max <= g.percent_dense * extent
    sel_clone = torch.logical_and(torch.where(torch.norm(grads, dim=-1) >= max_grad, True, False), small)      # :418-420
    # :398-402 -- the padded gradient of the rows appended by the clone step is zero: never selected for max_grad > 0
    sel_split = torch.logical_and(torch.where(grads.squeeze(-1) >= max_grad, True, False), ~small)
    # final prune_points (:440-445) on the densified model; max_radii2D was reset to zeros by densification_postfix
    low_opacity = (g.get_opacity < min_opacity).squeeze(-1)
    prune_self = low_opacity
    prune_child = low_opacity
    if max_screen_size:
        big_vs = torch.zeros_like(scale_max) > max_screen_size
        prune_self = torch.logical_or(torch.logical_or(low_opacity, big_vs), scale_max > 0.1 * extent)
        child_scaling = g.scaling_inverse_activation(g.get_scaling / (0.8 * N))            # :410
        child_max = torch.max(g.scaling_activation(child_scaling), dim=1).values
        prune_child = torch.logical_or(torch.logical_or(low_opacity, big_vs), child_max > 0.1 * extent)
    keep = torch.logical_and(~sel_split, ~prune_self)
    clone = torch.logical_and(sel_clone, ~prune_self)
    child = torch.logical_and(sel_split, ~prune_child)
    return keep, clone, child, sel_split


def densify_and_prune(g, max_grad, min_opacity, extent, max_screen_size, N=2, _timing=None):
    """`_timing` (a dict, development only): filled with wall-clock milliseconds per phase, each closed by a device
    synchronisation."""
    import time

    def _mark(name, _t=[None]):
        if _timing is None:
            return
        torch.cuda.synchronize()
        now = time.perf_counter()
        if _t[0] is not None:
            _timing[name] = round((now - _t[0]) * 1e3, 3)
        _t[0] = now
    if max_grad <= 0:
        raise RuntimeError("bags_densify: max_grad must be positive (with a non-positive threshold the reference also "
                           "splits the rows it has just cloned)")
    xyz = g._xyz
    if not xyz.is_cuda:
        raise RuntimeError("bags_b200: densification needs CUDA tensors (there is no CPU path)")
    dev = xyz.device
    P = xyz.shape[0]
    with torch.no_grad(), torch.cuda.device(dev):
        _mark("start")
        keep, clone, child, split = selection_masks(g, max_grad, min_opacity, extent, max_screen_size, N)
        _mark("masks")
        # four scans along the contiguous axis (a cumsum down the rows of a [P, 4] tensor is ~100x slower), then
        # interleaved to the [P][4] records the kernel reads with one 16-byte load
        bits = torch.stack([keep, clone, child, split], dim=0).to(torch.int32)              # [4, P]
        incl = torch.cumsum(bits, dim=1, dtype=torch.int32)
        plan = (incl - bits).t().contiguous()                                               # exclusive scans, [P, 4]
        flags = (bits[0] | (bits[1] << 1) | (bits[2] << 2) | (bits[3] << 3)).to(torch.uint8).contiguous()
        kept, clones, children, parents = (incl[:, -1].tolist() if P > 0 else [0, 0, 0, 0])  # the one host sync
        _mark("scan")
        # same random draw as densify_and_split (:404-406)
        stds = g.get_scaling[split].repeat(N, 1)
        means = torch.zeros((stds.size(0), 3), device=dev)
        samples = torch.normal(mean=means, std=stds).contiguous() if parents > 0 else means
        P_out = kept + clones + N * children
        rotation = g._rotation.detach().contiguous()
        _mark("samples")

        descs, news = [], []
        for group in g.optimizer.param_groups:
            assert len(group["params"]) == 1
            old = group["params"][0]
            if old.shape[0] != P:
                raise RuntimeError(f"bags_densify: parameter group {group['name']} has {old.shape[0]} rows, expected {P}")
            src = old.detach().contiguous()
            width = src.numel() // max(P, 1)
            dst = torch.empty((P_out,) + tuple(src.shape[1:]), dtype=torch.float32, device=dev)
            state = g.optimizer.state.get(old, None)
            sm = sv = dm = dv = None
            if state is not None and "exp_avg" in state:
                sm, sv = state["exp_avg"].contiguous(), state["exp_avg_sq"].contiguous()
                dm, dv = torch.empty_like(dst), torch.empty_like(dst)
            descs.append(capi.DensifyTensor(src.data_ptr(), dst.data_ptr(), sm.data_ptr() if sm is not None else None,
                                            dm.data_ptr() if dm is not None else None,
                                            sv.data_ptr() if sv is not None else None,
                                            dv.data_ptr() if dv is not None else None, int(width),
                                            ROLE.get(group["name"], 0)))
            news.append((group, old, state, dst, dm, dv, (src, sm, sv)))
        _mark("alloc")
        arr = (capi.DensifyTensor * len(descs))(*descs)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = capi.lib().bags_densify_gather(P, N, plan.data_ptr(), flags.data_ptr(), kept, clones, children, parents,
                                            rotation.data_ptr(), samples.data_ptr(), len(descs), arr, stream)
        capi.check(rc, "bags_densify_gather")
        _mark("gather")

        # hand the new tensors to the optimizer the way cat_tensors_to_optimizer / _prune_optimizer do (:315-373)
        for group, old, state, dst, dm, dv, _keepalive in news:
            if state is not None:
                if dm is not None:
                    state["exp_avg"], state["exp_avg_sq"] = dm, dv
                del g.optimizer.state[old]
                group["params"][0] = nn.Parameter(dst.requires_grad_(True))
                g.optimizer.state[group["params"][0]] = state
            else:
                group["params"][0] = nn.Parameter(dst.requires_grad_(True))
            attr = GROUP_ATTR.get(group["name"])
            if attr is not None:
                setattr(g, attr, group["params"][0])
        g.xyz_gradient_accum = torch.zeros((P_out, 1), device=dev)                          # :390-393 then :345-348
        g.denom = torch.zeros((P_out, 1), device=dev)
        g.max_radii2D = torch.zeros((P_out,), device=dev)
        _mark("rebind")
    return {"kept": kept, "clones": clones, "children": N * children, "split_parents": parents, "rows": P_out}
