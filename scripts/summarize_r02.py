#!/usr/bin/env python
"""Turn the logs of scripts/r02_call{1,2,3}_*.sh (gpurun_out/r02_*.log) into one markdown table: per variant the pytest
verdict or the bench line's headline numbers (ms/step, value, e2e, roofline fraction, secondary steps/s).
  python scripts/summarize_r02.py [gpurun_out] > profiles/r02_variants.md"""
import glob
import json
import os
import re
import sys


def last_json(text):
    for line in reversed(text.strip().splitlines()):
        line = line.strip()
        if line.startswith('{') and line.endswith('}'):
            try:
                return json.loads(line)
            except ValueError:
                continue
    return None


def describe(path):
    text = open(path, errors='replace').read()
    name = os.path.basename(path)[len('r02_'):-len('.log')]
    verdict = re.findall(r'^(?:=+ )?((?:\d+ (?:passed|failed|skipped|deselected|error|errors|warnings?)(?:, )?)+) in [\d.]+s', text, re.M)
    if verdict:
        return name, 'pytest', verdict[-1]
    js = last_json(text)
    if js is None:
        tail = ' / '.join(text.strip().splitlines()[-2:])[:160]
        return name, 'log', tail or '(empty)'
    if 'metric' not in js:                                     # --only-secondary prints the bare dict
        return name, 'secondary', ', '.join(f'{k}={v:.4g}' for k, v in js.items() if isinstance(v, (int, float)))
    parts = [f"{js.get('ms_per_step', float('nan')):.2f} ms/step", f"{js.get('value', float('nan')):.4g} {js.get('unit', '')}",
             f"n_gpus={js.get('n_gpus')}"]
    if isinstance(js.get('e2e'), dict):
        parts.append(f"e2e {js['e2e'].get('value', float('nan')):.4g}")
    if isinstance(js.get('roofline'), dict):
        parts.append(f"agg frac {js['roofline'].get('frac', float('nan')):.3f} ({js['roofline'].get('avg_launch_ms', float('nan')):.2f} ms/pass)")
    cfg = js.get('config', {})
    for k in ('parallelism', 'edge_locality', 'degree_skew'):
        if cfg.get(k):
            parts.append(f'{k}={cfg[k]}')
    sec = js.get('secondary')
    if isinstance(sec, dict):
        parts.append('secondary: ' + ', '.join(f'{k}={v:.4g}' for k, v in sec.items() if isinstance(v, (int, float))))
    if isinstance(js.get('clocks'), dict) and js['clocks'].get('reasons'):
        parts.append('clocks: ' + ','.join(js['clocks']['reasons']))
    return name, 'bench', '; '.join(parts)


def main():
    root = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out'
    rows = [describe(p) for p in sorted(glob.glob(os.path.join(root, 'r02_*.log')))]
    print('| run | kind | result |\n|---|---|---|')
    for name, kind, what in rows:
        print(f'| `{name}` | {kind} | {what} |')
    for status in sorted(glob.glob(os.path.join(root, 'r02_call*_status.txt'))):
        print(f'\n`{os.path.basename(status)}`:\n```\n{open(status).read().strip()}\n```')


if __name__ == '__main__':
    main()
