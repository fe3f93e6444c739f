"""CPU, world_size 2 over gloo: the observation-sharding plumbing.  The kernels
need a GPU, so each rank's partial Hessian / gradient / IJ columns come from the
oracle's closed forms; what is tested is the product's sharding logic
(shard_range, the all-reduce helper, column gathering) and that the sharded
assembly reproduces the single-process answer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, d, seed, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from vittles_b200 import distributed as vd
    from oracle import models
    r, w, group = vd.init_from_env(backend='gloo')
    assert (r, w) == (rank, world) and group is not None
    r0, r1 = vd.shard_range(n, rank, world)
    X, y, _ = models.synth_logistic(seed, r1 - r0, d, row0=r0)        # each rank generates only its rows
    theta = np.linspace(-0.5, 0.5, d)
    cf = models.glm_closed_form(X, y, theta, np.ones(r1 - r0))
    H = torch.as_tensor(cf['hessian'].copy())
    g = torch.as_tensor(cf['grad'].copy())
    vd.allreduce_sum_(H, group)
    vd.allreduce_sum_(g, group)
    S_local = torch.as_tensor(-np.linalg.solve(H.numpy(), cf['cross_hessian']))   # this rank's columns only
    S_all = vd.gather_columns(S_local, group)
    if rank == 0:
        np.savez(os.path.join(out_dir, 'dist_out.npz'), H=H.numpy(), g=g.numpy(), S=S_all.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_assembly_matches_single_process(tmp_path):
    from oracle import models
    n, d, seed, world = 1001, 6, 5, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, d, seed, str(tmp_path)), nprocs=world, join=True)
    out = np.load(os.path.join(str(tmp_path), 'dist_out.npz'))
    X, y, _ = models.synth_logistic(seed, n, d)
    theta = np.linspace(-0.5, 0.5, d)
    cf = models.glm_closed_form(X, y, theta, np.ones(n))
    np.testing.assert_allclose(out['H'], cf['hessian'], rtol=1e-12)
    np.testing.assert_allclose(out['g'], cf['grad'], rtol=1e-10, atol=1e-12)
    S = -np.linalg.solve(cf['hessian'], cf['cross_hessian'])
    np.testing.assert_allclose(out['S'], S, rtol=1e-9, atol=1e-13)
