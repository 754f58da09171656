"""The formation of BASELINE config 3 solved by dual decomposition (the reference's
``FormationPoint2pointDualDecomposition``, omgtools/problems/formation_dualdec.py): every agent's NLP
holds its own trajectory and copies of its neighbours', one batched xz-update per iteration on the
GPU(s), dual ascent with step --rho.

    python examples/formation_dualdec.py --agents 8 --iters 40 --rho 0.02
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29541 examples/formation_dualdec.py --agents 64
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--agents', type=int, default=8)
    ap.add_argument('--iters', type=int, default=40)
    ap.add_argument('--rho', type=float, default=0.02)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('needs a CUDA device (no CPU fallback)')
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
    from omg_tools_b200.problems.admm_gpu import FormationDDRunner
    pr = sc.config_formation_dd(args.agents, options={'rho': args.rho}, rank=rank, world=world)
    run = FormationDDRunner(pr, rank=rank, world=world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = [run.dual_update(0.) for _ in range(args.iters)]
    e1.record()
    torch.cuda.synchronize()
    st, it = run.status()
    if rank == 0:
        print(json.dumps({'agents': args.agents, 'n_gpus': world, 'rho': args.rho, 'iterations': args.iters,
                          'ms_per_iteration': e0.elapsed_time(e1) / args.iters,
                          'primal_residual_first': res[0], 'primal_residual_min': min(res),
                          'primal_residual_last': res[-1], 'xz_updates_succeeded': bool((st == 0).all()),
                          'mean_ip_iterations_last': float(it.mean())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
