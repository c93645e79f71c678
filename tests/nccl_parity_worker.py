#!/usr/bin/env python
"""Multi-GPU parity worker (run under torchrun, one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/nccl_parity_worker.py

Every rank renders its share of `V` views of one scene through this repo's rasterizer (gradient sinks into the flat
buffer, balanced view assignment, the step captured in a CUDA graph on odd ranks and run eagerly on even ranks so
that both schedules meet in one reduction), then bags_view_parallel.GradientReducer sums the Gaussian gradients and
the densification statistics over NCCL.  Rank 0 recomputes all V views alone with the UNMODIFIED REFERENCE build
(oracle/_ref) and checks:
  * reduced flat gradient == sum over views of the reference's single-view gradients (1e-4 relative L2 per field),
  * reduced densification statistics / max_radii2D == the single-process accumulation of the reference loop
    (scene/gaussian_model.py:449-455, train.py:377),
  * the global shift_factors gradient rides along in the small collective,
  * per-camera pose gradients stayed on the owning rank and match the reference's.
Prints one JSON line on rank 0 and exits non-zero on any mismatch.  Used by tests/test_nccl_gpu.py.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import parity_common as pc  # noqa: E402
import bags_view_parallel as VP  # noqa: E402

S = pc.S


def render_view(mod, params, cam, leaves, shift, bg, sh_degree=3):
    view, proj, campos, intr = leaves
    st = mod.GaussianRasterizationSettings(
        image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
        viewmatrix=view, projmatrix=proj, intrinsic=intr, sh_degree=sh_degree, campos=campos, prefiltered=False,
        debug=False, debug_iter=0)
    m2d = torch.zeros_like(params["means3D"], requires_grad=True)
    m2dd = torch.zeros_like(params["means3D"], requires_grad=True)
    out = mod.GaussianRasterizer(st)(means3D=params["means3D"], means2D=m2d, means2D_densify=m2dd, shift_factors=shift,
                                     opacities=params["opacities"], shs=params["shs"], scales=params["scales"],
                                     rotations=params["rotations"])
    return out, m2d, m2dd


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")     # the overlapped slice all-reduces must not queue behind compute
    dist.init_process_group("nccl", device_id=dev)
    mod = pc.load_ours()
    W, H, P, V = 640, 368, 200_000, 3 * world
    cam0 = S.make_camera(W, H, seed=100)
    gs = S.make_gaussians(P, cam0, seed=0, scale=0.01)
    cams = [S.make_camera(W, H, seed=100 + v) for v in range(V)]
    ups = [[u.to(dev) for u in S.make_upstream(H, W, seed=1 + v)] for v in range(V)]
    bg = torch.zeros(3, device=dev)
    shift0 = torch.tensor([1e-3, -5e-4, 2e-4])

    def fresh_leaves(v):
        c = cams[v]
        return (c.viewmatrix.to(dev).requires_grad_(True), c.projmatrix.to(dev).requires_grad_(True),
                c.campos.to(dev).requires_grad_(True), c.intrinsic.to(dev))

    # ---- ours, view-parallel ------------------------------------------------------------------------------------------
    params = VP.FlatGaussianParams(gs.means3D.to(dev), gs.shs.to(dev), gs.opacities.to(dev), gs.scales.to(dev),
                                   gs.rotations.to(dev), device=dev)
    params.enable_direct_accumulation(mod)
    shift = shift0.to(dev).requires_grad_(True)
    # measured work of the round-robin share -> balanced assignment (every rank computes the same table)
    local_costs = []
    for v in VP.shard_views(V, rank, world):
        with torch.no_grad():
            (c, *_), _, _ = render_view(mod, params, cams[v], fresh_leaves(v), shift, bg)
        local_costs.append(float((c != 0).sum()))          # any deterministic per-view number works as a cost
    costs = VP.gather_view_costs(local_costs, V, rank, world)
    table = VP.balance_views(costs, world)
    mine = table[rank]
    leaves = {v: fresh_leaves(v) for v in mine}
    stats = VP.DensificationStats(P, dev)
    reducer = VP.P2PGradientReducer(params)      # the peer-memory kernel (falls back to NCCL if the buffers cannot be mapped)
    m2d_grads = {}

    def body():
        params.zero_grad()
        stats.zero()
        with params.accumulating():
            for v in mine:
                (color, radii, depth, weights, _), m2d, m2dd = render_view(mod, params, cams[v], leaves[v], shift, bg)
                torch.autograd.backward([color, depth, weights], ups[v])
                stats.add_view(m2d.grad, m2dd.grad, radii)
                m2d_grads[v] = m2d.grad

    def clear():
        shift.grad = None
        for v in mine:
            for t in leaves[v][:3]:
                t.grad = None

    clear()
    if rank % 2 == 1:
        step = VP.CapturedStep(body, dev, mod, capacity=6_000_000, forwards_per_step=len(mine), reset=clear)
        step.replay()
        torch.cuda.synchronize()
        step.check()
        schedule = "captured graph"
    else:
        body()
        schedule = "eager"
    reducer.reduce()
    reducer.reduce_stats(stats, shift.grad)
    torch.cuda.synchronize()
    flat_per_view = params.flat_grad.clone()
    shift_grad_per_view = shift.grad.clone()

    # ---- ours, the rank's views as ONE launch set, gradient exchange overlapped slice by slice ----------------------------
    # (GaussianBatchRasterizer + SlicedGradientReducer; run eagerly, then captured in a CUDA graph -- NCCL calls included)
    sliced = VP.SlicedGradientReducer(params)
    settings = [mod.GaussianRasterizationSettings(
        image_height=H, image_width=W, tanfovx=cams[v].tanfovx, tanfovy=cams[v].tanfovy, bg=bg, scale_modifier=1.0,
        viewmatrix=leaves[v][0], projmatrix=leaves[v][1], intrinsic=leaves[v][3], sh_degree=3, campos=leaves[v][2],
        prefiltered=False, debug=False, debug_iter=0) for v in mine]
    brast = mod.GaussianBatchRasterizer(settings, 6_000_000, backward_slices=3, on_slice=sliced.on_slice)
    bm2d = torch.zeros(len(mine), P, 3, device=dev, requires_grad=True)
    bm2dd = torch.zeros(len(mine), P, 3, device=dev, requires_grad=True)
    bups = [torch.stack([ups[v][k] for v in mine]).contiguous() for k in range(3)]
    bshift = shift0.to(dev).requires_grad_(True)

    def batch_body():
        params.zero_grad()
        with params.accumulating():
            color, radii, depth, weights, _ = brast(means3D=params["means3D"], means2D=bm2d, means2D_densify=bm2dd,
                                                    shift_factors=bshift, opacities=params["opacities"],
                                                    shs=params["shs"], scales=params["scales"],
                                                    rotations=params["rotations"])
            torch.autograd.backward([color, depth, weights], bups)
        sliced.finish()

    def batch_clear():
        clear()
        bm2d.grad = bm2dd.grad = bshift.grad = None

    # every rank takes the same path here: the collectives are part of the body, so the eager warm-up runs of the capture
    # and the replay must happen on all ranks alike
    batch_clear()
    batch_body()                                  # eager
    torch.cuda.synchronize()
    brast.check()
    flat_batch_eager = params.flat_grad.clone()
    batch_clear()
    bstep = VP.CapturedStep(batch_body, dev, mod, None, forwards_per_step=1, reset=batch_clear, check=brast.check)
    bstep.replay()
    torch.cuda.synchronize()
    bstep.check()
    batch_graph_ok = pc.rel_l2(params.flat_grad.cpu(), flat_batch_eager.cpu()) <= 1e-5
    flat_batch = params.flat_grad.clone()
    batch_pose_local = {v: torch.cat([leaves[v][0].grad.flatten(), leaves[v][1].grad.flatten(),
                                      leaves[v][2].grad]).cpu() for v in mine}
    batch_m2d_ok = batch_graph_ok and all(pc.rel_l2(bm2d.grad[k].cpu(), m2d_grads[v].cpu()) <= 1e-5
                                          for k, v in enumerate(mine))

    # ---- gather the per-camera pose gradients for the check on rank 0 --------------------------------------------------
    pose_local = {v: torch.cat([leaves[v][0].grad.flatten(), leaves[v][1].grad.flatten(), leaves[v][2].grad]).cpu()
                  for v in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, (mine, pose_local, schedule, batch_pose_local, batch_m2d_ok))

    ok, report = True, {}
    if rank == 0:
        ref = pc.load_ref()
        rparams = VP.FlatGaussianParams(gs.means3D.to(dev), gs.shs.to(dev), gs.opacities.to(dev), gs.scales.to(dev),
                                        gs.rotations.to(dev), device=dev)
        rshift = shift0.to(dev).requires_grad_(True)
        rstats = VP.DensificationStats(P, dev)
        rpose = {}
        rparams.zero_grad()
        for v in range(V):
            lv = fresh_leaves(v)
            (color, radii, depth, weights, _), m2d, m2dd = render_view(ref, rparams, cams[v], lv, rshift, bg)
            torch.autograd.backward([color, depth, weights], ups[v])
            rstats.add_view(m2d.grad, m2dd.grad, radii)
            rpose[v] = torch.cat([lv[0].grad.flatten(), lv[1].grad.flatten(), lv[2].grad]).cpu()
        torch.cuda.synchronize()
        off = 0
        for name, shape in VP.FIELDS:
            n = P
            for s_ in shape:
                n *= s_
            e = pc.rel_l2(flat_per_view[off:off + n].cpu(), rparams.flat_grad[off:off + n].cpu())
            report["grad_" + name] = e
            ok &= e <= 1e-4
            e = pc.rel_l2(flat_batch[off:off + n].cpu(), rparams.flat_grad[off:off + n].cpu())
            report["batch_grad_" + name] = e
            ok &= e <= 1e-4
            off += n
        for k in ("accum", "accum_abs", "denom"):
            e = pc.rel_l2(getattr(stats, k).cpu(), getattr(rstats, k).cpu())
            report["stats_" + k] = e
            ok &= e <= 1e-4
        report["max_radii2D_mismatch"] = int((stats.max_radii2D != rstats.max_radii2D).sum())
        ok &= report["max_radii2D_mismatch"] == 0
        report["shift_grad"] = pc.rel_l2(shift_grad_per_view.cpu(), rshift.grad.cpu())
        ok &= report["shift_grad"] <= 1e-4
        worst = 0.0
        owners = {}
        worst_b = 0.0
        for r_, (views_r, pose_r, sched, bpose_r, bm2d_ok) in enumerate(gathered):
            ok &= bool(bm2d_ok)
            for v in views_r:
                owners[v] = (r_, sched)
                # view matrix block [0:16] is the reference's racy term (Q13): compare proj + campos
                worst = max(worst, pc.rel_l2(pose_r[v][16:], rpose[v][16:]))
                worst_b = max(worst_b, pc.rel_l2(bpose_r[v][16:], rpose[v][16:]))
        report["pose_proj_campos_worst"] = worst
        report["batch_pose_proj_campos_worst"] = worst_b
        ok &= worst <= 1e-4 and worst_b <= 1e-4 and sorted(owners) == list(range(V))
        report["assignment"] = table
        report["schedules"] = [g[2] for g in gathered]
        report["world"] = world
        report["gradient_exchange"] = reducer.describe()
        report["nccl"] = ".".join(str(x) for x in torch.cuda.nccl.version())
        report["ok"] = bool(ok)
        print(json.dumps(report), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    params.disable_direct_accumulation()
    rc = 0 if int(flag) == 1 else 1
    torch.cuda.synchronize()
    del bstep                                  # the captured graph holds NCCL work: release it before the group goes
    dist.barrier()
    sys.stdout.flush()
    os._exit(rc)                               # (destroy_process_group after graph-captured collectives was seen to hang)


if __name__ == "__main__":
    sys.exit(main())
