"""Array plumbing: the reference works on numpy arrays; this build computes on
float64 CUDA tensors.  Inputs may be numpy arrays, python sequences or torch
tensors on any device; results are returned in the kind of the input they
correspond to (numpy in -> numpy out), so existing numpy code keeps working,
while CUDA tensors stay on the device end to end."""
import numpy as np
import torch


def default_device():
    if not torch.cuda.is_available():
        raise RuntimeError('vittles_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.')
    return torch.device('cuda', torch.cuda.current_device())


def to_device(a, device=None):
    """float64 CUDA tensor from numpy / sequence / tensor (no copy if already there)."""
    if device is None:
        device = default_device()
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64)
    return torch.as_tensor(np.asarray(a, dtype=np.float64), device=device)


def kind_of(a):
    """'numpy', 'cpu' or 'cuda' - how a result matching `a` should be returned."""
    if isinstance(a, torch.Tensor):
        return 'cuda' if a.is_cuda else 'cpu'
    return 'numpy'


PINNED_STAGING_MIN_BYTES = 64 << 20


def to_host(t):
    """CPU copy of a device tensor.  Large results (the (D, N) sensitivity matrix is 82 GB at N = 1e7) go through a
    PINNED buffer with one asynchronous copy - pageable memory moves at a fraction of the PCIe rate; the buffer comes
    from torch's caching host allocator, so repeated calls reuse it."""
    if not t.is_cuda:
        return t
    t = t.detach()
    if t.numel() * t.element_size() < PINNED_STAGING_MIN_BYTES:
        return t.cpu()
    src = t if t.is_contiguous() else t.contiguous()
    # torch's caching host allocator hands out pinned blocks in powers of two (an 82 GB result would lock 128 GB):
    # when that block would take most of what the host has left, go through the small staging buffers instead -
    # an over-committed pinned allocation does not raise, it gets the process killed
    nbytes = src.numel() * src.element_size()
    avail = _host_available_bytes()
    if avail is not None and (1 << max(nbytes - 1, 1).bit_length()) > 0.7 * avail:
        return _to_host_staged(src)
    try:
        out = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
    except RuntimeError:
        # the host cannot pin that much (the 82 GB result next to 82 GB of pinned inputs): pageable destination,
        # filled through two pinned staging buffers - the CPU copy of chunk i-1 runs behind the transfer of chunk i
        return _to_host_staged(src)
    out.copy_(src, non_blocking=True)
    torch.cuda.current_stream(src.device).synchronize()
    return out


STAGING_BYTES = 512 << 20


def _host_available_bytes():
    try:
        import psutil
        return int(psutil.virtual_memory().available)
    except Exception:
        return None


def _to_host_staged(src):
    flat = src.reshape(-1)
    n = flat.numel()
    avail = _host_available_bytes()
    if avail is not None and n * src.element_size() > 0.9 * avail:
        raise MemoryError('the result ({:.1f} GB) does not fit in the available host memory ({:.1f} GB); keep it on the '
                          'device (pass CUDA tensors in, get CUDA tensors back)'.format(n * src.element_size() / 1e9, avail / 1e9))
    out = torch.empty(n, dtype=src.dtype)
    per = max(1, STAGING_BYTES // src.element_size())
    stage = [torch.empty(min(per, n), dtype=src.dtype, pin_memory=True) for _ in range(2)]
    events = [torch.cuda.Event(), torch.cuda.Event()]
    stream = torch.cuda.current_stream(src.device)
    pending = None                                   # (offset, length, buffer index) of the chunk in flight
    for i, off in enumerate(range(0, n, per)):
        b = i & 1
        ln = min(per, n - off)
        stage[b][:ln].copy_(flat[off:off + ln], non_blocking=True)
        events[b].record(stream)
        if pending is not None:
            po, pl, pb = pending
            events[pb].synchronize()
            out[po:po + pl].copy_(stage[pb][:pl])
        pending = (off, ln, b)
    if pending is not None:
        po, pl, pb = pending
        events[pb].synchronize()
        out[po:po + pl].copy_(stage[pb][:pl])
    return out.reshape(src.shape)


def as_kind(t, kind):
    if kind == 'cuda':
        return t
    if kind == 'cpu':
        return to_host(t)
    return to_host(t).numpy()


def like(t, proto):
    return as_kind(t, kind_of(proto))
