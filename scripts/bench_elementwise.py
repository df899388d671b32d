#!/usr/bin/env python
"""Times the BatchNorm(+ReLU+dropout) kernels and the column statistics at the C5 shapes, first versions against the
column-stationary variants (GNNB200_EW_V2=1, csrc/elementwise_v2.cu), and checks that the variants reproduce the
first versions (bitwise for the elementwise kernels, to rounding for the statistics).

  python scripts/bench_elementwise.py [--rows 2449029] [--reps 20]

The flag is read once per process, so the script re-runs itself as a child with the variable set."""
import argparse
import json
import os
import subprocess
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
    res = {}
    for name, flag in (('v1', '0'), ('v2', '1')):
        env = dict(os.environ, GNNB200_EW_V2=flag)
        p = subprocess.run([sys.executable, __file__, '--child', '--rows', str(args.rows), '--reps', str(args.reps)],
                           env=env, capture_output=True, text=True)
        if p.returncode != 0:
            print(f'{name} failed:\n{p.stderr[-2000:]}')
            sys.exit(1)
        res[name] = json.loads(p.stdout.strip().splitlines()[-1])
    print(f'{"case":<18}{"kernel":<20}{"v1 ms":>9}{"frac":>7}{"v2 ms":>9}{"frac":>7}')
    ok = True
    for case in res['v1']:
        for k, a in res['v1'][case].items():
            if k == 'checks':
                for ck, va in a.items():
                    vb = res['v2'][case]['checks'][ck]
                    exact = case.endswith('colstats') is False and ck in ('y', 'y_absmax', 'dx', 'dx_absmax')
                    same = (va == vb) if exact else abs(va - vb) <= 1e-5 * max(1.0, abs(va))
                    ok &= same
                    if not same:
                        print(f'  MISMATCH {case}.{ck}: v1 {va!r} v2 {vb!r}')
                continue
            b = res['v2'][case][k]
            print(f'{case:<18}{k:<20}{a["ms"]:9.3f}{a["frac_of_hbm_peak"]:7.2f}{b["ms"]:9.3f}{b["frac_of_hbm_peak"]:7.2f}')
    print('variants reproduce the first versions' if ok else 'VARIANTS DISAGREE')
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
