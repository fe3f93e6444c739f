"""Factorise and solve block-arrow systems on the GPU (the B200 replacement for
"SuperLU on the COO matrix", ``solver_lib.py:46-48``).

    H = [ blockdiag(B_g)   C ]   ->   L_g = chol(B_g),  Z_g = L_g^{-1} C_g,
        [ C^T            Hgg ]        S = Hgg - Z^T Z = chol-factorised dense Schur complement

    solve(b):  y_g = L_g^{-1} b_g ;  x_glob = S^{-1} (b_glob - Z^T y) ;
               x_g = L_g^{-T} (y_g - Z_g x_glob)
"""
import torch

from . import ops
from ._arrays import to_device, kind_of, as_kind


BLOCK_MAXM = 32          # csrc/blockchol.cuh: one block per warp / CTA out of shared memory


class BlockArrowSolver:
    def __init__(self, h, overwrite=False):
        self.h = h
        self.d = h.shape[0]
        self.sa = h.sparsity_array
        self.gi = h.global_inds
        self.G, self.M = self.sa.shape
        self.Dg = int(self.gi.numel())
        if h.blocks is None:
            raise ValueError('the block-diagonal part of the Hessian is missing')
        covered = self.G * self.M + self.Dg
        if covered != self.d:
            raise ValueError('block-arrow solver: local and global indices cover {} of {} parameters'.format(
                covered, self.d))
        blocks = h.blocks if overwrite else h.blocks.clone()
        self.Lb = ops.block_potrf(blocks.contiguous())
        self.Z = None
        self.schur = None
        if self.Dg > 0:
            if h.hgg is None:
                raise ValueError('the global block of the Hessian is missing')
            S = h.hgg.clone()
            if h.cross is not None:
                cross = h.cross if overwrite else h.cross.clone()
                self.Z = ops.block_trsm(self.Lb, cross.contiguous())
                Z2 = self.Z.reshape(self.G * self.M, self.Dg)
                S = S - ops.syrk_weighted(Z2)            # Z^T Z on the FP64 tensor-core engine
            self.schur = ops.potrf(S, overwrite=True)

    def _solve_vec(self, b):
        y = b[self.sa].contiguous()                      # (G, M) gather of the local right-hand sides
        ops.block_solve(self.Lb, y, transpose=False)
        x = torch.empty_like(b)
        if self.Dg > 0:
            rhs = b[self.gi].contiguous()
            if self.Z is not None:
                Z2 = self.Z.reshape(self.G * self.M, self.Dg)
                rhs = ops.tall_colsum(Z2, y.reshape(-1), alpha=-1.0, y0=rhs, beta=1.0)
            xg = self.schur.solve(rhs)
            x[self.gi] = xg
            if self.Z is not None:
                ops.tall_gemv(Z2, xg, alpha=-1.0, y=y.reshape(-1), beta=1.0)
        ops.block_solve(self.Lb, y, transpose=True)
        x[self.sa] = y
        return x

    def solve(self, v):
        kind = kind_of(v)
        b = to_device(v, self.Lb.device)
        if b.shape[0] != self.d or b.dim() > 2:
            raise ValueError('right-hand side has shape {}, expected ({},) or ({}, K)'.format(
                tuple(b.shape), self.d, self.d))
        if b.dim() == 1:
            return as_kind(self._solve_vec(b.contiguous()), kind)
        cols = [self._solve_vec(b[:, k].contiguous()) for k in range(b.shape[1])]
        return as_kind(torch.stack(cols, dim=1), kind)
