#!/usr/bin/env python3
"""Build tuning variants of the library and bench them in ONE gpurun call.

    tools/sweep.py base= s3="-DDT_STAGES_V=3" n12="-DDT_MCSA=12" [--steps 4] [--bench-args "--no-e2e"]

Every NAME=FLAGS pair is built with tools/build_variant.sh into build/variants/ (an empty FLAGS string is the default
build), then `bench.py --no-cpu-baseline --no-e2e` runs once per variant on the same box and the per-stage times are
printed side by side.  Differences below ~1 % are box-to-box / run-to-run noise: put a baseline variant in every sweep."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    steps, extra, variants = 4, '--no-e2e', []
    args = sys.argv[1:]
    while args:
        a = args.pop(0)
        if a == '--steps':
            steps = int(args.pop(0))
        elif a == '--bench-args':
            extra = args.pop(0)
        else:
            name, _, flags = a.partition('=')
            variants.append((name, flags))
    procs = [subprocess.Popen([os.path.join(ROOT, 'tools', 'build_variant.sh'), n] + f.split(), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for n, f in variants]
    for (n, _), p in zip(variants, procs):
        out = p.communicate()[0]
        if p.returncode != 0:
            sys.exit('build of %s failed:\n%s' % (n, out[-3000:]))
    names = ' '.join(n for n, _ in variants)
    cmd = ('for v in %s; do APPLES_B200_LIB=build/variants/libapples_b200_$v.so python bench.py --steps %d --warmup 3 '
           '--no-cpu-baseline %s 2>/dev/null > gpurun_out/sweep_$v.json; done' % (names, steps, extra))
    for n, _ in variants:
        try:
            os.remove(os.path.join(ROOT, 'gpurun_out', 'sweep_%s.json' % n))
        except OSError:
            pass
    r = subprocess.run(['/usr/local/graft/bin/gpurun', '--timeout', str(120 + 40 * len(variants)), '--', cmd], cwd=ROOT,
                       capture_output=True, text=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:])
    print('%-12s %9s %9s %9s %9s' % ('variant', 'ms/step', 'dense', 'select', 'place'))
    for n, _ in variants:
        try:
            d = json.loads(open(os.path.join(ROOT, 'gpurun_out', 'sweep_%s.json' % n)).read().strip().splitlines()[-1])
            s = d['stage_ms_per_step']
            print('%-12s %9.2f %9.2f %9.2f %9.2f' % (n, d['ms_per_step'], s['rep_distance_ms'], s['selection_ms'], s['placement_ms']))
        except Exception as e:  # noqa: BLE001
            print('%-12s failed (%s)' % (n, e))


if __name__ == '__main__':
    main()
