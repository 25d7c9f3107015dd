#!/usr/bin/env python
"""Where the end-to-end step spends its time: the e2e leg of bench.py with stages removed and with 1..12 stream slots.
Run under gpurun: python tools/e2e_breakdown.py [atoms]"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from arpeggio_b200 import params, synth  # noqa: E402
from arpeggio_b200.batch import BatchRunner  # noqa: E402
from arpeggio_b200.engine import pinned_soa  # noqa: E402

atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
p = params.make_params()
soa = synth.cloud_featured(atoms, seed=2, h_decimals=3)
hosts = {'plain': pinned_soa(soa), 'wire': pinned_soa(soa.to_wire())}
steps = 240


def staged(runner, host, mode):
    """mode: 'all' upload + run + packed fetch; 'no_upload' run + fetch; 'no_fetch' upload + run + count; 'kernels' run + count;
    'upload' upload only (+ sync); 'fetch' fetch of a finished run is not separable (the sort is redone per run)"""
    todo = list(range(steps))
    lock = threading.Lock()

    def work(slot):
        eng = runner.engines[slot]
        eng.upload_atoms(host, check_finite=False)
        n0 = eng.run_pairs()
        buf = runner._packed_buffer(slot, host.n_atoms, n0 + 1024, False)
        while True:
            with lock:
                if not todo:
                    return
                todo.pop()
            if mode in ('all', 'no_fetch', 'upload'):
                eng.upload_atoms(host, check_finite=False)
            if mode == 'upload':
                eng.sync()
                continue
            eng.run_pairs_async()
            if mode in ('all', 'no_upload'):
                eng.fetch_pairs_packed(False, out=buf)
            else:
                eng.pair_count()

    ths = [threading.Thread(target=work, args=(s,)) for s in range(len(runner.engines))]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for e in runner.engines:
        e.sync()
    return (time.perf_counter() - t0) / steps * 1e3


def utilisation(fn):
    """GPU utilisation (share of time with a kernel running, nvidia-smi's sampling) while fn() runs repeatedly for about 2 s"""
    import subprocess
    pr = subprocess.Popen(['nvidia-smi', '--query-gpu=utilization.gpu', '--format=csv,noheader,nounits', '-lms', '100', '-i', '0'],
                          stdout=subprocess.PIPE, text=True)
    t0 = time.perf_counter()
    ms = []
    while time.perf_counter() - t0 < 2.5:
        ms.append(fn())
    pr.terminate()
    vals = [int(v) for v in pr.communicate()[0].split() if v.strip().isdigit()]
    return round(float(np.median(ms)), 4), vals


if os.environ.get('UTIL'):
    with BatchRunner(device=0, slots=6, params=p) as runner:
        for mode in ('kernels', 'no_upload', 'all'):
            for name, host in hosts.items():
                print(mode, name, utilisation(lambda: staged(runner, host, mode)), flush=True)
    sys.exit(0)

for slots in (1, 2, 3, 6, 12):
    with BatchRunner(device=0, slots=slots, params=p) as runner:
        for name, host in hosts.items():
            staged(runner, host, 'all')
            row = {m: round(staged(runner, host, m), 4) for m in ('all', 'no_upload', 'no_fetch', 'kernels', 'upload')}
            print(f'slots={slots:2d} {name:5s} h2d={host.input_bytes() / 1e6:.2f} MB  ms/step: {row}', flush=True)
