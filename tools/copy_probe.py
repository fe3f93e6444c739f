import time, torch
dev = torch.device('cuda', 0)
N, D = 1_000_000, 1024
Xh = torch.empty((N, D), dtype=torch.float64, pin_memory=True); Xh.fill_(1.0)
Xd = torch.empty((N, D), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
cs = torch.cuda.Stream(device=dev)
def t(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print('%-48s host %.1f ms, total %.1f ms' % (label, (t1 - t0) * 1e3, (t2 - t0) * 1e3), flush=True)
def whole_default(): Xd.copy_(Xh, non_blocking=True)
def chunks_side():
    with torch.cuda.stream(cs):
        for c in range(16):
            r0, r1 = N * c // 16, N * (c + 1) // 16
            Xd[r0:r1].copy_(Xh[r0:r1], non_blocking=True)
def chunks_side_events():
    evs = []
    cs.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(cs):
        for c in range(16):
            r0, r1 = N * c // 16, N * (c + 1) // 16
            Xd[r0:r1].copy_(Xh[r0:r1], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(cs); evs.append(ev)
    return evs
def chunks_then_small_blocking():
    chunks_side()
    y = torch.ones(1000, dtype=torch.float64).to(dev)     # pageable -> blocking copy on the current stream
def chunks_then_small_pinned():
    chunks_side()
    y = torch.ones(1000, dtype=torch.float64).pin_memory().to(dev, non_blocking=True)
t('whole tensor, default stream, non_blocking', whole_default)
t('16 chunks on a side stream', chunks_side)
t('16 chunks + events', chunks_side_events)
t('16 chunks, then small pageable .to(dev)', chunks_then_small_blocking)
t('16 chunks, then small pinned non_blocking', chunks_then_small_pinned)
