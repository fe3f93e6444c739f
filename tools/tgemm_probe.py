#!/usr/bin/env python3
"""Staged bring-up / timing probe of the tcgen05 TF32 engine (csrc/tgemm.cu).

    python tools/tgemm_probe.py            # every stage, each in its own process (a device trap poisons a context)
    python tools/tgemm_probe.py STAGE      # one stage in this process
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ['convert', 'kc_one', 'kc_big', 'kc_ragged', 'kc_x3', 'ks_one', 'ks_big', 'ks_x3', 'splitk', 'apply', 'syrk',
          'time_apply', 'time_syrk', 'time_parts']


def tf32_round(x32):
    import torch
    b = x32.view(torch.int32)
    return ((b + 0x1000) & ~0x1FFF).view(torch.float32)


def err_report(name, out, ref, extra=None):
    import torch
    d = (out - ref).abs()
    scale = float(ref.abs().max())
    res = {'stage': name, 'max_abs_err': float(d.max()), 'ref_scale': scale, 'rel': float(d.max()) / max(scale, 1e-300),
           'nan': bool(torch.isnan(out).any())}
    if res['rel'] > 1e-2 or res['nan']:
        # where is it wrong?  error per 32-row band and per 32-column band
        M, N = d.shape
        rb = [float(d[i:i + 32].max()) for i in range(0, min(M, 256), 32)]
        cb = [float(d[:, j:j + 32].max()) for j in range(0, min(N, 512), 32)]
        res['row_bands'] = [round(v, 4) for v in rb]
        res['col_bands'] = [round(v, 4) for v in cb]
        res['out_00'] = [float(v) for v in out[0, :4]]
        res['ref_00'] = [float(v) for v in ref[0, :4]]
    if extra:
        res.update(extra)
    print(json.dumps(res), flush=True)
    return res


def run(stage):
    import torch
    from vittles_b200 import ops
    dev = torch.device('cuda', 0)
    g = torch.Generator(device=dev).manual_seed(1)

    def rnd(*shape):
        return torch.randn(*shape, device=dev, dtype=torch.float64, generator=g)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    if stage == 'convert':
        X = rnd(300, 102)
        hi, lo = ops.tf32_convert(X, split=3)
        ref_hi = tf32_round(X.float())
        ok_hi = bool(torch.equal(hi[:, :102], ref_hi)) and bool((hi[:, 102:] == 0).all())
        rem = (X - hi[:, :102].double()).abs().max() / X.abs().max()
        rem2 = (X - hi[:, :102].double() - lo[:, :102].double()).abs().max() / X.abs().max()
        print(json.dumps({'stage': stage, 'hi_exact': ok_hi, 'hi_rel_resid': float(rem), 'hi_lo_rel_resid': float(rem2)}))
        return

    def gemm_case(name, M, N, K, mode, precision, **kw):
        A = rnd(M, K) if mode == 'KC' else rnd(K, M)
        B = rnd(N, K) if mode == 'KC' else rnd(K, N)
        out = ops.tf32_gemm(A, B, mode, mode, precision=precision, **kw)
        torch.cuda.synchronize()
        Ah, Al = ops.tf32_convert(A, split=3)
        Bh, Bl = ops.tf32_convert(B, split=3)
        ca, cb = A.shape[1], B.shape[1]
        if precision == 'tf32':
            a64, b64 = Ah[:, :ca].double(), Bh[:, :cb].double()      # what the tensor core multiplies
        else:
            a64, b64 = A, B
        ref = a64 @ b64.T if mode == 'KC' else a64.T @ b64
        exact = A @ B.T if mode == 'KC' else A.T @ B
        return err_report(name, out, ref, {'rel_vs_f64': float((out - exact).abs().max() / exact.abs().max()),
                                           'shape': [M, N, K], 'mode': mode, 'precision': precision})

    if stage == 'kc_one':
        gemm_case(stage, 128, 256, 32, 'KC', 'tf32')
        gemm_case(stage + '_k64', 128, 256, 64, 'KC', 'tf32')
    elif stage == 'kc_big':
        gemm_case(stage, 1024, 2048, 1024, 'KC', 'tf32')
    elif stage == 'kc_ragged':
        gemm_case(stage, 200, 300, 100, 'KC', 'tf32')
        gemm_case(stage + '_k102', 130, 515, 102, 'KC', 'tf32')
    elif stage == 'kc_x3':
        gemm_case(stage, 1024, 2048, 1024, 'KC', 'tf32x3')
        gemm_case(stage + '_ragged', 200, 300, 100, 'KC', 'tf32x3')
    elif stage == 'ks_one':
        gemm_case(stage, 128, 256, 32, 'KS', 'tf32')
        gemm_case(stage + '_k64', 128, 256, 64, 'KS', 'tf32')
    elif stage == 'ks_big':
        gemm_case(stage, 1024, 1024, 4096, 'KS', 'tf32')
        gemm_case(stage + '_ragged', 200, 300, 1000, 'KS', 'tf32')
    elif stage == 'ks_x3':
        gemm_case(stage, 1024, 1024, 4096, 'KS', 'tf32x3')
    elif stage == 'splitk':
        gemm_case(stage + '_kc', 256, 256, 40000, 'KC', 'tf32')
        gemm_case(stage + '_ks', 256, 512, 100000, 'KS', 'tf32x3')
    elif stage in ('apply', 'time_apply'):
        D = 1024
        N = 20000 if stage == 'apply' else 1000000
        X = ops.synth_design(7, 0, N, D, dev)
        Hm = rnd(D, D)
        Hinv = (Hm @ Hm.T / D + torch.eye(D, device=dev, dtype=torch.float64))
        resid = rnd(N)
        ref = ops.ij_apply(Hinv, X, resid)
        for prec in ('tf32', 'tf32x3'):
            out = ops.ij_apply(Hinv, X, resid, precision=prec)
            torch.cuda.synchronize()
            extra = {'precision': prec, 'N': N}
            if stage == 'time_apply':
                ms = timed(lambda: ops.ij_apply(Hinv, X, resid, out=out, precision=prec))
                ms64 = timed(lambda: ops.ij_apply(Hinv, X, resid, out=ref), reps=1)
                extra.update({'ms': ms, 'tflops': 2.0 * D * D * N / ms / 1e9, 'ms_f64': ms64})
                ref = ops.ij_apply(Hinv, X, resid, out=ref)
            err_report(stage, out, ref, extra)
    elif stage in ('syrk', 'time_syrk'):
        D = 1024
        N = 50000 if stage == 'syrk' else 1000000
        X = ops.synth_design(7, 0, N, D, dev)
        s = torch.rand(N, device=dev, dtype=torch.float64, generator=g) * 0.25
        ref = ops.syrk_weighted(X, s)
        for prec in ('tf32', 'tf32x3'):
            out = ops.syrk_weighted(X, s, precision=prec)
            torch.cuda.synchronize()
            extra = {'precision': prec, 'N': N, 'symmetric': bool(torch.equal(out, out.T))}
            if stage == 'time_syrk':
                ms = timed(lambda: ops.syrk_weighted(X, s, out=out, precision=prec))
                ms64 = timed(lambda: ops.syrk_weighted(X, s, out=ref), reps=1)
                extra.update({'ms': ms, 'tflops_algorithmic': float(D) * (D + 1) * N / ms / 1e9, 'ms_f64': ms64})
            err_report(stage, out, ref, extra)
    elif stage == 'time_parts':
        # the pieces of the chunked apply, timed alone: conversion of one chunk, GEMM on pre-converted operands
        from vittles_b200 import _cabi
        from vittles_b200._cabi import ptr, stream, check
        lib = _cabi.require_cuda()
        D = 1024
        for split, prec in ((1, 'tf32'), (3, 'tf32x3')):
            for nchunk in (9472, 37888, 151552):
                X = ops.synth_design(7, 0, nchunk, D, dev)
                Hinv = rnd(D, D)
                ms_cvt = timed(lambda: ops.tf32_convert(X, split=split))
                Xh, Xl = ops.tf32_convert(X, split=split)
                Hh, Hl = ops.tf32_convert(Hinv, split=split)
                out = torch.empty((D, nchunk), dtype=torch.float64, device=dev)

                def gemm_only():
                    check(lib.vt_tf32_gemm(D, nchunk, D, -1.0, ptr(Hh), ptr(Hl), D, 0, ptr(Xh), ptr(Xl), D, 0, ptr(out),
                                           nchunk, None, None, None, 0, stream()))
                ms_gemm = timed(gemm_only, reps=5)
                print(json.dumps({'stage': stage, 'precision': prec, 'chunk_rows': nchunk, 'ms_convert': ms_cvt,
                                  'convert_gb_per_s': nchunk * D * (8 + 4 * (2 if split == 3 else 1)) / ms_cvt / 1e6,
                                  'ms_gemm': ms_gemm, 'gemm_tflops': 2.0 * D * D * nchunk / ms_gemm / 1e9,
                                  'gemm_out_gb_per_s': 8.0 * D * nchunk / ms_gemm / 1e6}), flush=True)
    else:
        raise SystemExit('unknown stage ' + stage)


if __name__ == '__main__':
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for st in STAGES:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), st], capture_output=True, text=True,
                                   timeout=120)
                sys.stdout.write(r.stdout)
                if r.returncode != 0:
                    print(json.dumps({'stage': st, 'rc': r.returncode, 'stderr': r.stderr[-600:]}), flush=True)
            except subprocess.TimeoutExpired:
                print(json.dumps({'stage': st, 'timeout': True}), flush=True)
