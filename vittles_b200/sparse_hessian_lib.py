"""Block-sparse Hessians - the drop-in for ``vittles/sparse_hessian_lib.py``.

The reference assembles a ``scipy.sparse.coo_matrix`` entry by entry in Python
loops (``sparse_hessian_lib.py:98-102,137-153``) and factorises it with SuperLU
(``solver_lib.py:46-48``).  Here the same Hessian is kept on the GPU in its
natural block-arrow layout

    H = [ blockdiag(B_1 .. B_G)   C ]      B_g : (M, M) local blocks
        [ C^T                   Hgg ]      C_g : (M, Dg) local x global, Hgg : (Dg, Dg)

(:class:`BlockArrowHessian`), which converts to a scipy COO / dense matrix on
request and is what ``solver_lib.get_cholesky_solver`` factorises with the
batched block-Cholesky + Schur-complement kernels.

Return-type change with respect to the reference: ``get_block_hessian`` /
``get_global_hessian`` / ``get_hessian`` return a :class:`BlockArrowHessian`, not a
``scipy.sparse.coo_matrix`` (a COO matrix of 4e8 entries on the host is exactly
what this build avoids).  ``.tocoo()`` gives the reference's type, ``.todense()`` /
``.toarray()`` / ``@`` work as on a scipy matrix, and ``scipy.sparse.issparse`` is
False - ``solver_lib.get_cholesky_solver`` dispatches on the class itself.
"""
import numpy as np
import scipy.sparse
import torch
from torch import func as tf

from . import ops
from ._arrays import to_device, kind_of, as_kind
from .objectives import StructuredObjective


def detect_block_arrow(h, max_block=32):
    """Host-side structure detection for :meth:`BlockArrowHessian.from_sparse`.

    Global indices are the rows that touch more than half of the columns; the
    remaining rows (at least as many as the global ones) must fall into two or
    more connected components of one common size ``M <= max_block`` that are
    coupled only to themselves and to the global rows.  Returns ``(sparsity_array (G, M), global_inds (Dg,), blocks (G, M, M),
    cross (G, M, Dg) | None, hgg (Dg, Dg) | None)`` as numpy arrays, or ``None``.
    Duplicate COO entries are summed, as scipy does (the reference relies on
    that, ``sparse_hessian_lib.py:147-153``)."""
    from scipy.sparse import csgraph
    a = scipy.sparse.csr_matrix(h)
    a.sum_duplicates()
    d = a.shape[0]
    if a.shape[0] != a.shape[1] or d < 2:
        return None
    pattern = (a != 0).astype(np.int8)
    pattern = ((pattern + pattern.T) > 0).astype(np.int8).tocsr()
    degree = np.diff(pattern.indptr)
    is_global = degree > max(d // 2, max_block)
    gi = np.nonzero(is_global)[0]
    li = np.nonzero(~is_global)[0]
    if len(li) < len(gi) or len(li) == 0:                    # mostly dense: not an arrow worth exploiting
        return None
    ncomp, labels = csgraph.connected_components(pattern[li][:, li], directed=False)
    sizes = np.bincount(labels, minlength=ncomp)
    M = int(sizes[0])
    if ncomp < 2 or M > max_block or np.any(sizes != M):
        return None
    order = np.argsort(labels, kind='stable')                # component-major, indices ascending inside
    sa = li[order].reshape(ncomp, M)
    G, Dg = ncomp, len(gi)
    perm = sa.reshape(-1)
    loc = a[perm][:, perm].tocoo()
    if np.any(loc.row // M != loc.col // M):
        return None
    blocks = np.zeros((G, M, M))
    blocks[loc.row // M, loc.row % M, loc.col % M] = loc.data
    cross = hgg = None
    if Dg > 0:
        cross = np.asarray(a[perm][:, gi].todense()).reshape(G, M, Dg)
        hgg = np.asarray(a[gi][:, gi].todense())
    return sa, gi, blocks, cross, hgg


class BlockArrowHessian:
    """Device-resident block-arrow matrix of dimension ``d``.

    ``blocks`` (G, M, M) | ``cross`` (G, M, Dg) | ``hgg`` (Dg, Dg); any part may
    be ``None`` (= zero).  ``sparsity_array`` (G, M) and ``global_inds`` (Dg,)
    place the parts in the flat parameter vector."""

    def __init__(self, d, sparsity_array, global_inds, blocks=None, cross=None, hgg=None, group=None):
        self.shape = (d, d)
        self.sparsity_array = sparsity_array      # torch int64 (G, M) on the device
        self.global_inds = global_inds            # torch int64 (Dg,)
        self.blocks, self.cross, self.hgg = blocks, cross, hgg
        # torch.distributed group over which the blocks are SHARDED (every rank holds the blocks / cross blocks of
        # its own observations; `hgg` is the complete global block, replicated); None: everything is here
        self.group = group

    @classmethod
    def from_sparse(cls, h, max_block=32, device=None):
        """Recognise block-arrow structure in a scipy sparse matrix (e.g. the
        ``coo_matrix`` the reference's ``SparseBlockHessian`` returns,
        ``sparse_hessian_lib.py:107,162``) and move it to the GPU - SURVEY.md
        section 8f item 3.  ``None`` when the pattern does not fit
        (:func:`detect_block_arrow`); the caller then densifies."""
        parts = detect_block_arrow(h, max_block=max_block)
        if parts is None:
            return None
        sa, gi, blocks, cross, hgg = parts
        dev = to_device(np.zeros(1), device).device
        return cls(h.shape[0], torch.as_tensor(sa, dtype=torch.int64, device=dev),
                   torch.as_tensor(gi, dtype=torch.int64, device=dev), blocks=to_device(blocks, dev),
                   cross=None if cross is None else to_device(cross, dev),
                   hgg=None if hgg is None else to_device(hgg, dev))

    def __add__(self, other):
        if not isinstance(other, BlockArrowHessian) or other.shape != self.shape:
            return NotImplemented

        def add(a, b):
            return b if a is None else (a if b is None else a + b)
        ginds = self.global_inds if self.global_inds.numel() >= other.global_inds.numel() else other.global_inds
        return BlockArrowHessian(self.shape[0], self.sparsity_array, ginds, add(self.blocks, other.blocks),
                                 add(self.cross, other.cross), add(self.hgg, other.hgg),
                                 group=self.group if self.group is not None else other.group)

    # -- conversions (small problems / interoperability) ----------------------
    def tocoo(self):
        rows, cols, vals = [], [], []
        sa = self.sparsity_array
        G, M = sa.shape
        if self.blocks is not None:
            rows.append(sa[:, :, None].expand(G, M, M).reshape(-1))
            cols.append(sa[:, None, :].expand(G, M, M).reshape(-1))
            vals.append(self.blocks.reshape(-1))
        gi = self.global_inds
        Dg = gi.numel()
        if self.cross is not None and Dg > 0:
            r = sa[:, :, None].expand(G, M, Dg).reshape(-1)
            c = gi[None, None, :].expand(G, M, Dg).reshape(-1)
            v = self.cross.reshape(-1)
            rows += [r, c]; cols += [c, r]; vals += [v, v]
        if self.hgg is not None and Dg > 0:
            rows.append(gi[:, None].expand(Dg, Dg).reshape(-1))
            cols.append(gi[None, :].expand(Dg, Dg).reshape(-1))
            vals.append(self.hgg.reshape(-1))
        if not vals:
            return scipy.sparse.coo_matrix(self.shape)
        r = torch.cat(rows).cpu().numpy()
        c = torch.cat(cols).cpu().numpy()
        v = torch.cat(vals).cpu().numpy()
        return scipy.sparse.coo_matrix((v, (r, c)), self.shape)

    def todense(self):
        return np.asarray(self.tocoo().todense())

    def to_dense_tensor(self):
        """The (d, d) matrix as a float64 tensor on the device (small problems only)."""
        d = self.shape[0]
        sa, gi = self.sparsity_array, self.global_inds
        G, M = sa.shape
        out = torch.zeros((d, d), dtype=torch.float64, device=sa.device)
        if self.blocks is not None:
            out.index_put_((sa[:, :, None].expand(G, M, M), sa[:, None, :].expand(G, M, M)), self.blocks, accumulate=True)
        Dg = gi.numel()
        if self.cross is not None and Dg > 0:
            r = sa[:, :, None].expand(G, M, Dg)
            c = gi[None, None, :].expand(G, M, Dg)
            out.index_put_((r, c), self.cross, accumulate=True)
            out.index_put_((c, r), self.cross, accumulate=True)
        if self.hgg is not None and Dg > 0:
            out.index_put_((gi[:, None].expand(Dg, Dg), gi[None, :].expand(Dg, Dg)), self.hgg, accumulate=True)
        return out

    def toarray(self):
        return self.todense()

    def matvec(self, v):
        """``H @ v`` for a (d,) or (d, K) array, without densifying."""
        kind = kind_of(v)
        x = to_device(v, self.sparsity_array.device)
        vec = x.dim() == 1
        x2 = x.reshape(self.shape[0], -1)
        sa, gi = self.sparsity_array, self.global_inds
        out = torch.zeros_like(x2)
        xl = x2[sa]                                        # (G, M, K)
        if self.blocks is not None:
            out[sa] += torch.einsum('gij,gjk->gik', self.blocks, xl)
        if gi.numel() > 0:
            xg = x2[gi]                                    # (Dg, K)
            if self.cross is not None:
                out[sa] += torch.einsum('gij,jk->gik', self.cross, xg)
                out[gi] += torch.einsum('gij,gik->jk', self.cross, xl)
            if self.hgg is not None:
                out[gi] += self.hgg @ xg
        return as_kind(out.reshape(-1) if vec else out, kind)

    def __matmul__(self, v):
        return self.matvec(v)

    def dot(self, v):
        return self.matvec(v)

    def get_solver(self):
        """``solve(v) -> H^{-1} v`` by batched block Cholesky and a dense Schur
        complement on the global block."""
        from .block_solver import BlockArrowSolver, BLOCK_MAXM
        if self.sparsity_array.shape[1] > BLOCK_MAXM:
            # the batched kernels keep one block per warp / CTA in shared memory (M <= 32); wider blocks go the way
            # the reference's SuperLU goes - a general factorisation - here the dense GPU Cholesky
            if self.group is not None:
                raise ValueError('block-arrow solver: blocks wider than {} are not supported on a sharded Hessian'
                                 .format(BLOCK_MAXM))
            if self.shape[0] > 32768:
                raise ValueError('block-arrow solver: blocks of size {} exceed the batched kernels\' limit of {} and '
                                 'the matrix (dimension {}) is too large to densify'.format(
                                     self.sparsity_array.shape[1], BLOCK_MAXM, self.shape[0]))
            from .solver_lib import get_dense_cholesky_solver
            return get_dense_cholesky_solver(self.to_dense_tensor())
        return BlockArrowSolver(self).solve


class SparseBlockHessian():
    """Block-diagonal (plus dense global rows) Hessian of
    ``f(x) = sum_g f_g(x_g, x_global)`` (reference: ``sparse_hessian_lib.py:11-168``).

    ``objective_function`` is a torch callable of the flat parameter;
    ``sparsity_array`` (G, M) lists the indices of every block (unique, equal
    block sizes - ``ValueError`` otherwise, ``:55-57``).  Like the reference,
    the block Hessian costs M Hessian-vector products (one per within-block
    index, ``:62-67``) and the global rows one per global index (``:128-136``);
    the scatter into blocks is a single device gather instead of the
    reference's G*M Python iterations."""

    def __init__(self, objective_function, sparsity_array):
        self._fun = objective_function
        sa = np.asarray(sparsity_array.cpu() if isinstance(sparsity_array, torch.Tensor) else sparsity_array)
        if sa.ndim != 2:
            raise ValueError('``sparsity_array`` must be (num_blocks, block_size).')
        if len(np.unique(sa)) != sa.size:
            raise ValueError('The indices in ``sparsity array`` must be unique.')
        self._sparsity_array = sa
        self._num_blocks, self._block_size = sa.shape
        self._structured = isinstance(objective_function, StructuredObjective)
        if not self._structured:
            self._f_grad = tf.grad(self._fun)

    def _hvp(self, x, v):
        return tf.jvp(self._f_grad, (x,), (v,))[1]

    def _prep(self, opt_par):
        x = to_device(opt_par)
        x = torch.atleast_1d(x)
        if x.dim() != 1:
            raise ValueError('``opt_par`` must be a vector.')
        sa = torch.as_tensor(self._sparsity_array, dtype=torch.int64, device=x.device)
        return x, sa

    def get_block_hessian(self, opt_par, print_every=0):
        """Reference ``:69-108``."""
        x, sa = self._prep(opt_par)
        if self._structured:
            return self._fun.vt_block_hessian(x, sa, which='block')
        G, M = sa.shape
        blocks = torch.empty((G, M, M), dtype=torch.float64, device=x.device)
        for ib in range(M):
            if print_every > 0 and ib % print_every == 0:
                print('Block index {} of {}.'.format(ib, M))
            v = torch.zeros_like(x)
            v[sa[:, ib]] = 1
            blocks[:, :, ib] = self._hvp(x, v)[sa]
        if print_every > 0:
            print('Done differentiating.')
        empty = torch.empty(0, dtype=torch.int64, device=x.device)
        return BlockArrowHessian(len(x), sa, empty, blocks=blocks)

    def get_global_hessian(self, opt_par, global_inds=None, print_every=0):
        """Reference ``:110-163``: dense rows/columns of the global parameters
        (by default every index not in ``sparsity_array``); global and local
        indices must be disjoint (``ValueError``, ``:118-122``)."""
        x, sa = self._prep(opt_par)
        local_inds = self._sparsity_array.reshape(-1)
        if global_inds is None:
            global_inds = np.setdiff1d(np.arange(len(x)), local_inds)
        global_inds = np.asarray(global_inds).reshape(-1)
        inter = np.intersect1d(global_inds, local_inds)
        if len(inter) > 0:
            raise ValueError('The global and local indices must be disjoint.  {}'.format(inter))
        gi = torch.as_tensor(global_inds, dtype=torch.int64, device=x.device)
        if self._structured:
            return self._fun.vt_block_hessian(x, sa, which='global', global_inds=gi)
        G, M = sa.shape
        Dg = len(global_inds)
        cross = torch.empty((G, M, Dg), dtype=torch.float64, device=x.device)
        hgg = torch.empty((Dg, Dg), dtype=torch.float64, device=x.device)
        for j in range(Dg):
            if print_every > 0 and j % print_every == 0:
                print('Global index {} of {}.'.format(j, Dg))
            v = torch.zeros_like(x)
            v[gi[j]] = 1
            hrow = self._hvp(x, v)
            cross[:, :, j] = hrow[sa]
            hgg[:, j] = hrow[gi]
        if print_every > 0:
            print('Done differentiating.')
        # the reference writes every global x global entry as two halves, one from
        # each end (:146-153): the assembled block is the symmetrised one
        hgg = 0.5 * (hgg + hgg.T)
        return BlockArrowHessian(len(x), sa, gi, cross=cross, hgg=hgg)

    def get_hessian(self, opt_par, print_every=0):
        """Reference ``:165-168``."""
        if self._structured:
            x, sa = self._prep(opt_par)
            return self._fun.vt_block_hessian(x, sa, which='full')
        return self.get_block_hessian(opt_par, print_every=print_every) + \
            self.get_global_hessian(opt_par, print_every=print_every)
