"""Oracle restatement of ``vittles/sparse_hessian_lib.py``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).
"""
import numpy as np
import torch
from torch import func as tf
from scipy.sparse import coo_matrix


def _hvp_fun(f):
    """forward-over-reverse Hessian-vector product, as
    ``_append_jvp(autograd.grad(f))`` (``sparse_hessian_lib.py:59-60``)."""
    g = tf.grad(f)

    def hvp(x, v):
        xt = torch.as_tensor(np.asarray(x, dtype=np.float64))
        vt = torch.as_tensor(np.asarray(v, dtype=np.float64))
        return tf.jvp(g, (xt,), (vt,))[1].detach().numpy().copy()
    return hvp


def check_sparsity_array(sparsity_array):
    """``sparse_hessian_lib.py:55-57``."""
    sparsity_array = np.asarray(sparsity_array)
    if len(np.unique(sparsity_array)) != sparsity_array.size:
        raise ValueError('The indices in ``sparsity array`` must be unique.')
    return sparsity_array


def block_hessian(f, opt_par, sparsity_array):
    """``get_block_hessian`` (``sparse_hessian_lib.py:69-108``): one HVP per
    within-block index ``ib`` against the indicator of ``sparsity_array[:, ib]``
    (``:62-67``), scattered so that block ``b`` receives column
    ``sparsity_array[b, ib]`` (``:98-102``)."""
    sparsity_array = check_sparsity_array(sparsity_array)
    opt_par = np.atleast_1d(np.asarray(opt_par, dtype=np.float64))
    if opt_par.ndim != 1:
        raise ValueError('``opt_par`` must be a vector.')
    hvp = _hvp_fun(f)
    G, M = sparsity_array.shape
    vals, rows, cols = [], [], []
    for ib in range(M):
        v = np.zeros_like(opt_par)
        v[sparsity_array[:, ib]] = 1
        hp = hvp(opt_par, v)
        vals.append(hp[sparsity_array].reshape(-1))                  # (G*M,) rows of every block
        rows.append(sparsity_array.reshape(-1))
        cols.append(np.repeat(sparsity_array[:, ib], M))
    d = len(opt_par)
    return coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), (d, d))


def global_hessian(f, opt_par, sparsity_array, global_inds=None):
    """``get_global_hessian`` (``sparse_hessian_lib.py:110-163``): one HVP per
    global index; local x global entries written twice (``:137-144``),
    global x global entries written as two halves so that COO duplicate
    summation restores them (``:146-153``)."""
    sparsity_array = np.asarray(sparsity_array)
    opt_par = np.asarray(opt_par, dtype=np.float64)
    local_inds = np.hstack(sparsity_array)
    if global_inds is None:
        global_inds = np.setdiff1d(np.arange(len(opt_par)), local_inds)
    global_inds = np.asarray(global_inds)
    inter = np.intersect1d(global_inds, local_inds)
    if len(inter) > 0:
        raise ValueError('The global and local indices must be disjoint.  {}'.format(inter))
    hvp = _hvp_fun(f)
    vals, rows, cols = [], [], []
    for ig in global_inds:
        v = np.zeros_like(opt_par)
        v[ig] = 1
        hr = hvp(opt_par, v)
        ig_l = np.full(len(local_inds), ig)
        vals += [hr[local_inds], hr[local_inds]]
        rows += [local_inds, ig_l]
        cols += [ig_l, local_inds]
        ig_g = np.full(len(global_inds), ig)
        vals += [0.5 * hr[global_inds], 0.5 * hr[global_inds]]
        rows += [global_inds, ig_g]
        cols += [ig_g, global_inds]
    d = len(opt_par)
    if not vals:
        return coo_matrix((d, d))
    return coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), (d, d))


def full_hessian(f, opt_par, sparsity_array):
    """``get_hessian`` (``sparse_hessian_lib.py:165-168``)."""
    return block_hessian(f, opt_par, sparsity_array) + global_hessian(f, opt_par, sparsity_array)
