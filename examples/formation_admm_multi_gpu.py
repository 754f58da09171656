"""BASELINE config 3: FormationPoint2point ADMM, 64 holonomic agents on a ring,
2 rectangular obstacles, agents sharded over the GPUs of one box; the consensus
exchange (x_i, then z_ij / l_ij) and the residual sum go over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
        --master-addr 127.0.0.1 --master-port 29531 examples/formation_admm_multi_gpu.py --agents 64

With --check every rank-sharded quantity is compared against a single-GPU run of
all agents on rank 0 (same kernels, no exchange).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=64)
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--check', action='store_true')
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    os.environ['OMG_B200_DEVICE'] = str(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    from omg_tools_b200 import scenarios as sc
    from omg_tools_b200.problems.admm_gpu import FormationADMMRunner
    pr = sc.config3(args.agents)
    run = FormationADMMRunner(pr, rank=rank, world=world)
    for _ in range(3):                       # warm-up iterations (also ADMM init_iter)
        run.dual_update(0.)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.iters):
        res = run.dual_update(0.)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    tm = torch.tensor([dt], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    st, it = run.status()
    ok = torch.tensor([float((st == 0).all())], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    out = {'workload': 'config3: FormationPoint2point ADMM, %d agents, ring, rho=1' % args.agents,
           'n_gpus': world, 'admm_iterations_per_s': args.iters / float(tm[0]),
           'agent_x_updates_per_s': args.agents * args.iters / float(tm[0]),
           'primal_res': res[0], 'dual_res': res[1], 'all_x_updates_succeeded': bool(ok[0] > 0)}
    if args.check:
        # gather the sharded state and compare with a single-GPU run of all agents
        xs = [torch.empty_like(run.x_i) for _ in range(world)] if world > 1 else [run.x_i]
        zs = [torch.empty_like(run.z_i) for _ in range(world)] if world > 1 else [run.z_i]
        if world > 1:
            dist.all_gather(xs, run.x_i.contiguous())
            dist.all_gather(zs, run.z_i.contiguous())
        if rank == 0:
            ref = FormationADMMRunner(sc.config3(args.agents), rank=0, world=1)
            for _ in range(3 + args.iters):
                ref.dual_update(0.)
            out['max_abs_dx_vs_single_gpu'] = float((torch.cat(xs) - ref.x_i).abs().max())
            out['max_abs_dz_vs_single_gpu'] = float((torch.cat(zs) - ref.z_i).abs().max())
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
