#!/usr/bin/env python
"""Times the BatchNorm(+ReLU+dropout) kernels and the column statistics at the C5 shapes (L2 flushed between launches).
The round-2 comparison of the first versions against the column-stationary kernels that are now the default is kept in
profiles/r02/a_elementwise_v1_v2.txt.

  python scripts/bench_elementwise.py [--rows 2449029] [--reps 20]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def measure(rows, reps):
    import torch
    import gnnb200  # noqa: F401
    from gnnb200 import ops
    dev = torch.device('cuda')
    peak = 6550.4
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        peak = float(json.load(open(path))['hbm_gbs'])
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)
    for cols in (256, 512):
        x = torch.randn(rows, cols, device=dev, generator=g) * 2 + 1
        gy = torch.randn(rows, cols, device=dev, generator=g)
        gamma = torch.rand(cols, device=dev, generator=g) + 0.5
        beta = torch.randn(cols, device=dev, generator=g) * 0.1
        flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)          # > L2

        def timed(fn, nbytes):
            fn()
            torch.cuda.synchronize()
            ms = []
            for _ in range(reps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ms.append(a.elapsed_time(b))
            ms.sort()
            t = ms[len(ms) // 2]
            return {'ms': t, 'GBs': nbytes / t / 1e6, 'frac_of_hbm_peak': nbytes / t / 1e6 / peak}

        n = rows * cols * 4
        s, m2 = ops.colstats.fn(x)
        mean, invstd = ops.bn_batch_stats.fn(x, None, None, 0.1, 1e-5)
        for drop in (0.0, 0.2):
            tag = f'c{cols}_p{drop}'
            y = ops.bn_act.fn(x, mean, invstd, gamma, beta, True, drop, 1234, True, 0)
            dx, dg, db = ops.bn_act_bwd.fn(gy, x, mean, invstd, gamma, beta, True, drop, 1234, True)
            out[tag] = {
                'fwd': timed(lambda: ops.bn_act.fn(x, mean, invstd, gamma, beta, True, drop, 1234, True, 0), 2 * n),
                'bwd_reduce+apply': timed(lambda: ops.bn_act_bwd.fn(gy, x, mean, invstd, gamma, beta, True, drop, 1234, True), 5 * n),
                'checks': {'y': float(y.double().sum()), 'y_absmax': float(y.abs().max()), 'dx': float(dx.double().sum()),
                           'dx_absmax': float(dx.abs().max()), 'dgamma': float(dg.double().sum()), 'dbeta': float(db.double().sum())},
            }
            del y, dx
        out[f'c{cols}_colstats'] = {'colstats': timed(lambda: ops.colstats.fn(x), n),
                                    'checks': {'sum': float(s.double().sum()), 'm2': float(m2.double().sum())}}
        del x, gy
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--rows', type=int, default=2_449_029)
    ap.add_argument('--reps', type=int, default=20)
    ap.add_argument('--child', action='store_true')
    args = ap.parse_args()
    if args.child:
        print(json.dumps(measure(args.rows, args.reps)))
        return
    res = measure(args.rows, args.reps)
    print(f'{"case":<18}{"kernel":<20}{"ms":>9}{"frac of HBM peak":>18}')
    for case, entry in res.items():
        for k, a in entry.items():
            if k != 'checks':
                print(f'{case:<18}{k:<20}{a["ms"]:9.3f}{a["frac_of_hbm_peak"]:18.2f}')


if __name__ == '__main__':
    main()
