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


# ---------------------------------------------------------------------------
# Sharded block-arrow solve (BASELINE config 3 on N > 1 GPUs): every rank holds the blocks of its own
# observations, the Dg x Dg Schur complement is all-reduced once at factor time and the Dg x K reduced right-hand
# side once per solve.  The kernels need a GPU, so the ranks plug a torch-CPU stand-in for
# the kernel set into the product's BlockArrowSolver; what is tested is its sharding logic.
# ---------------------------------------------------------------------------

class TorchCpuBlockKernels:
    @staticmethod
    def block_potrf(blocks):
        return torch.linalg.cholesky(blocks)

    @staticmethod
    def block_trsm(Lb, C, transpose=False):
        A = Lb.transpose(1, 2) if transpose else Lb
        C.copy_(torch.linalg.solve_triangular(A, C, upper=transpose))
        return C

    @staticmethod
    def block_solve(Lb, y, transpose=False):
        A = Lb.transpose(1, 2) if transpose else Lb
        y.copy_(torch.linalg.solve_triangular(A, y.unsqueeze(-1), upper=transpose).squeeze(-1))
        return y

    @staticmethod
    def gram(Z2):
        return Z2.T @ Z2

    @staticmethod
    def dense_factor(S):
        class F:
            L = torch.linalg.cholesky(S)

            def solve(self, rhs):
                return torch.cholesky_solve(rhs.reshape(rhs.shape[0], -1), self.L).reshape(rhs.shape)
        return F()

    @staticmethod
    def zt_times(Z2, Y2):
        return Z2.T @ Y2

    @staticmethod
    def sub_z_times(Z2, Xg, Y2):
        Y2 -= Z2 @ Xg
        return Y2


def _arrow_problem(G, M, Dg, K, seed):
    rng = np.random.RandomState(seed)
    blocks = rng.normal(size=(G, M, M))
    blocks = blocks @ blocks.transpose(0, 2, 1) + 3.0 * np.eye(M)
    cross = 0.3 * rng.normal(size=(G, M, Dg))
    hgg = rng.normal(size=(Dg, Dg))
    hgg = hgg @ hgg.T + (5.0 + G) * np.eye(Dg)
    b_glob = rng.normal(size=(Dg, K))
    b_loc = rng.normal(size=(G, M, K))
    return blocks, cross, hgg, b_glob, b_loc


def _arrow_worker(rank, world, port, G, M, Dg, K, seed, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from vittles_b200 import distributed as vd
    from vittles_b200.sparse_hessian_lib import BlockArrowHessian
    from vittles_b200.block_solver import BlockArrowSolver
    _, _, group = vd.init_from_env(backend='gloo')
    blocks, cross, hgg, b_glob, b_loc = _arrow_problem(G, M, Dg, K, seed)
    g0, g1 = vd.shard_range(G, rank, world)
    gl = g1 - g0
    d_loc = Dg + gl * M                                       # (global replica, this rank's local parameters)
    sa = torch.as_tensor(Dg + np.arange(gl * M).reshape(gl, M))
    gi = torch.arange(Dg)
    h = BlockArrowHessian(d_loc, sa, gi, blocks=torch.as_tensor(blocks[g0:g1].copy()),
                          cross=torch.as_tensor(cross[g0:g1].copy()), hgg=torch.as_tensor(hgg.copy()), group=group)
    solver = BlockArrowSolver(h, kernels=TorchCpuBlockKernels)
    b = torch.as_tensor(np.concatenate([b_glob, b_loc[g0:g1].reshape(gl * M, K)], axis=0))
    x = solver.solve(b)
    x1 = solver.solve(b[:, 0].clone())                        # vector right-hand side
    assert tuple(x.shape) == (d_loc, K) and tuple(x1.shape) == (d_loc,)
    np.savez(os.path.join(out_dir, 'arrow_{}.npz'.format(rank)), x=x.numpy(), x1=x1.numpy(), g0=g0, g1=g1)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_block_arrow_solve_matches_dense(tmp_path):
    G, M, Dg, K, seed, world = 11, 3, 4, 5, 7, 2
    port = _free_port()
    mp.spawn(_arrow_worker, args=(world, port, G, M, Dg, K, seed, str(tmp_path)), nprocs=world, join=True)
    blocks, cross, hgg, b_glob, b_loc = _arrow_problem(G, M, Dg, K, seed)
    d = Dg + G * M
    H = np.zeros((d, d))
    H[:Dg, :Dg] = hgg
    for g in range(G):
        r = slice(Dg + g * M, Dg + (g + 1) * M)
        H[r, r] = blocks[g]
        H[r, :Dg] = cross[g]
        H[:Dg, r] = cross[g].T
    ref = np.linalg.solve(H, np.concatenate([b_glob, b_loc.reshape(G * M, K)], axis=0))
    for rank in range(world):
        out = np.load(os.path.join(str(tmp_path), 'arrow_{}.npz'.format(rank)))
        g0, g1 = int(out['g0']), int(out['g1'])
        np.testing.assert_allclose(out['x'][:Dg], ref[:Dg], rtol=1e-10, atol=1e-13)          # replicated global part
        np.testing.assert_allclose(out['x'][Dg:], ref[Dg + g0 * M:Dg + g1 * M], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(out['x1'][:Dg], ref[:Dg, 0], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(out['x1'][Dg:], ref[Dg + g0 * M:Dg + g1 * M, 0], rtol=1e-10, atol=1e-13)
