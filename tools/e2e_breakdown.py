#!/usr/bin/env python
"""Where the pipelined end-to-end step (8 streams, one host thread) spends its time: the step with stages removed.
    all        upload + kernels + sorted packed view + copies          (BatchRunner.run(packed=True))
    no_upload  inputs resident: kernels + view + copies
    no_d2h     upload + kernels + view, the words stay on the device   (ARPEGGIO_DEBUG_NO_D2H)
    device     inputs resident, words stay on the device: kernels + view only
    no_sort    upload + kernels, the count is the only result
    kernels    inputs resident: the pair kernels only
Run under gpurun: python tools/e2e_breakdown.py [atoms]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == 'child':
    from arpeggio_b200 import params, synth
    from arpeggio_b200.batch import BatchRunner
    from arpeggio_b200.engine import pinned_soa
    atoms, form = int(sys.argv[2]), sys.argv[3]
    p = params.make_params()
    soa = synth.cloud_featured(atoms, seed=2, h_decimals=3 if form == 'wire_h_fix' else None)
    host = pinned_soa(soa if form == 'plain' else soa.to_wire())
    steps, S = 480, 8

    def run(runner, upload, fetch):
        engs = runner.engines
        bufs = [runner._packed_buffer(s, host.n_atoms, 14 * host.n_atoms, False) for s in range(S)]
        pend = [False] * S
        t0 = time.perf_counter()
        for k in range(steps):
            s = k % S
            e = engs[s]
            if pend[s]:
                e.fetch_pairs_packed_wait() if fetch else e.pair_count()
            if upload:
                e.upload_atoms(host, check_finite=False)
            e.run_pairs_async()
            if fetch:
                e.fetch_pairs_packed_async(bufs[s], int(12.6 * host.n_atoms), False)
            pend[s] = True
        for s in range(S):
            if pend[s]:
                engs[s].fetch_pairs_packed_wait() if fetch else engs[s].pair_count()
            engs[s].sync()
        return (time.perf_counter() - t0) / steps * 1e3

    with BatchRunner(device=0, slots=S, params=p) as runner:
        for e in runner.engines:
            e.upload_atoms(host, check_finite=False)
            e.run_pairs()
        out = {}
        for name, up, fe in (('with upload, with view', True, True), ('resident, with view', False, True),
                             ('with upload, no view', True, False), ('resident, no view', False, False)):
            run(runner, up, fe)
            out[name] = round(min(run(runner, up, fe) for _ in range(3)), 4)
        print(out)
    sys.exit(0)

atoms = sys.argv[1] if len(sys.argv) > 1 else '100000'
for form in ('plain', 'wire', 'wire_h_fix'):
    for d2h in (True, False):
        env = dict(os.environ)
        if not d2h:
            env['ARPEGGIO_DEBUG_NO_D2H'] = '1'
        r = subprocess.run([sys.executable, __file__, 'child', atoms, form], env=env, capture_output=True, text=True)
        print(f'{form:10s} {"D2H of the words" if d2h else "words stay on the device":26s} ms/step: {r.stdout.strip() or r.stderr[-400:]}', flush=True)
