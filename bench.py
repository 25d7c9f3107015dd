#!/usr/bin/env python
"""Benchmark of the interatomic-contact hot path (BASELINE.json metric: classified atom-pairs/s on a
100k-atom synthetic structure).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--atoms A]

One step = one pass of the whole atom-atom job (cell-grid build + pair kernel: neighbour search,
filters, fused 15-bit CREDO classifier, record emission) over the configs[2] structure.
  value     pairs/s with the inputs resident in HBM; every step is bracketed by CUDA events on the
            library's stream and L2 is flushed (384 MiB memset) between steps, outside the brackets
  e2e       the same metric through the public API with HOST buffers: pinned H2D of the step's arrays +
            kernels + canonical (i, j) sort + D2H of the record stream inside the timed region.  The
            inputs cross PCIe in their wire form (soa.WireAtoms: uint8 counts for the CSR offsets, sparse
            halogen neighbours; decoded on the device), the stream in its packed form
            (arp_pairs_fetch_packed: row offsets + one 32-bit word per record; the float32 distances
            stay on the device until asked for); eight streams are driven by ONE host thread that
            enqueues every step whole (arp_pairs_run_async + arp_pairs_fetch_packed_async) and waits for it
            when its slot comes round again.  `legs` repeats the measurement with plain AtomSoA inputs,
            with the distance stream, with the compact 8-byte records, with the 16-byte records and with
            3-decimal hydrogens as int32 fixed point; e2e runs min(max(10 K, 100), 600) steps (its own
            `steps` key), the threaded legs K
  roofline  the step's algorithmic bytes (sum of input array bytes + 16 B per record) over the mean
            CUDA-event duration of the WHOLE step (grid build + pair kernels), against
            MEASURED_PEAKS.json (HBM copy); `pair_kernels` repeats it for the pair kernels alone and
            `large` for a 1M-atom cloud of the same recipe
  pcie      what the box's PCIe gives this rank while every rank copies at the same time: pinned
            cudaMemcpyAsync of the e2e step's sizes, H2D and D2H concurrently (arp_memcpy_probe)
  cpu_baseline / --impl reference
            the CPU oracle (oracle/arp_oracle.c, a C port of the reference's Python loop -- the
            reference itself needs BioPython/OpenBabel/gemmi, which are not installable here) on the
            host cores, one structure per thread
Multi-GPU (torchrun, one rank per GPU): structures are independent, so each rank runs its own
100k-atom structure (weak scaling, no collective on the data path; gloo only carries the barrier and
the max/sum of the per-rank results).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'classified atom-pairs/s'
UNIT = 'pairs/s'


def workload_name(atoms):
    return f'synthetic {atoms // 1000}k-atom cloud with SIFt feature masks, full 15-bit CREDO classifier (configs[2])'


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'MEASURED_PEAKS.json hbm_gbs (burst copy)'
    except Exception:
        return 6650.0, 'fallback 6.65 TB/s (B200_PROFILING.md)'


def ncu_traffic(atoms):
    """dram__bytes_read.sum + dram__bytes_write.sum of the step's kernels, per step, from the committed ncu --set full
    capture of this build (profiles/k_pairs_traffic.json, written by tools/make_profiles.py); None without one."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'k_pairs_traffic.json')) as fh:
            t = json.load(fh)
        if int(t.get('atoms', 0)) != atoms:
            return None, None
        return float(t['dram_bytes_per_launch']), (t.get('unflushed') or {}).get('l2_bytes_per_step')
    except Exception:
        return None, None


# ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={device}', f'--query-gpu={self.QUERY}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.15 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm = [float(r[1]) for r in rows if r[1].replace('.', '').isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': float(rows[0][2]) if rows[0][2].replace('.', '').isdigit() else None,
                'reasons': sorted(reasons), 'samples': len(rows)}


# ---------------------------------------------------------------------------------------------
class CpuPort:
    """The CPU oracle (C port of the reference loop) on `threads` host threads, one structure per thread
    (ctypes releases the GIL; the oracle keeps no global state).  A few distinct clouds are shared
    read-only by the threads."""

    def __init__(self, atoms_per_thread, threads, seed=2):
        from arpeggio_b200 import params as arp_params, synth
        from oracle import oracle
        self.oracle = oracle
        oracle.lib()
        self.p = arp_params.make_params()
        self.threads = threads
        distinct = [synth.cloud_featured(atoms_per_thread, seed=seed + 100 * t) for t in range(min(threads, 8))]
        self.soas = [distinct[t % len(distinct)] for t in range(threads)]

    def step(self, reps=1):
        """every thread classifies its structure `reps` times; returns (pairs, seconds)"""
        counts = [0] * self.threads

        def work(t):
            n = 0
            for _ in range(reps):
                n += self.oracle.pairs(self.soas[t], self.p).shape[0]
            counts[t] = n

        t0 = time.perf_counter()
        ths = [threading.Thread(target=work, args=(t,)) for t in range(self.threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        return sum(counts), time.perf_counter() - t0


def json_emitter_leg(records, n_atoms):
    """SURVEY 8 f2, host side: records -> the text of json.dump(get_contacts(), indent=4, sort_keys=True).
    C emitter (libarpeggio_cuda.so, arp_pairs_json_write) on all records against the reference's per-contact
    Python (get_contacts loop, interactions.py:183-196, + json.dumps) on a bounded sample of the same records."""
    from arpeggio_b200 import abi, jsonout
    atoms = [{'label_comp_id': 'ALA', 'auth_seq_id': a // 8, 'auth_asym_id': 'A', 'auth_atom_id': f'C{a % 8}',
              'pdbx_PDB_ins_code': ' ', 'label_comp_type': 'P'} for a in range(n_atoms)]
    rec = np.array(records)
    jsonout.pairs_json(rec[:1000], [jsonout.atom_fragment(d) for d in atoms])
    t0 = time.perf_counter()
    frags = [jsonout.atom_fragment(d) for d in atoms]          # one rendering per atom is part of the job
    text = jsonout.pairs_json(rec, frags)
    dt_c = time.perf_counter() - t0
    sample = rec[:20000]
    names = abi.SIFT_NAMES
    t0 = time.perf_counter()
    bag = []
    for r in sample:
        m = int(r['mask'])
        sifts = [m >> b & 1 for b in range(15)]
        e = {'bgn': dict(atoms[int(r['i'])]), 'end': dict(atoms[int(r['j'])]), 'type': 'atom-atom',
             'distance': round(np.float64(r['dist']), 2), 'contact': [k for k, v in zip(names, sifts) if v == 1],
             'interacting_entities': abi.CLASS_NAMES[(m >> 16) & 7]}
        bag.append(e)
    ref_text = json.dumps(bag, indent=4, sort_keys=True)
    dt_py = time.perf_counter() - t0
    ours = bytes(jsonout.pairs_json(sample, frags)).decode()
    assert '[\n' + ours + '\n]' == ref_text, 'emitter text differs from the Python dump'
    return {'emitter_pairs_per_s': rec.shape[0] / dt_c, 'emitter_bytes': len(text), 'emitter_s': dt_c,
            'emitter_threads': min(16, os.cpu_count() or 1),
            'python_pairs_per_s': sample.shape[0] / dt_py, 'python_sample': int(sample.shape[0]),
            'what': 'records -> contact JSON text (indent=4, sort_keys): C emitter on all records vs the per-contact Python '
                    'of get_contacts + json.dumps on a sample; texts compared byte for byte on the sample'}


def planes_leg_run(eng, soa, p, n_atoms, cpu=True):
    """configs[3]: the ring / amide plane terms on the synthetic plane set (2 048 rings, 12 500 amides, the configs[2]
    atoms): per term the wall time of one run + fetch through the public API (inputs resident), the record count
    and, beside it, the single-threaded CPU port of the reference's double loop."""
    from arpeggio_b200 import synth
    from oracle import oracle
    rings, amides = synth.plane_set(2048, 12_500, n_atoms=n_atoms, seed=3)
    eng.upload_atoms(soa)
    eng.upload_planes(rings, amides)
    out = {'rings': 2048, 'amides': 12_500, 'atoms': n_atoms, 'terms': {}}
    cpu_fn = {'ring_ring': lambda: oracle.ring_ring(rings, p), 'atom_ring': lambda: oracle.atom_ring(soa, rings, p),
              'amide_amide': lambda: oracle.amide_amide(amides, p), 'amide_ring': lambda: oracle.amide_ring(amides, rings, p)}
    got = eng.planes_all()
    t0 = time.perf_counter()
    for _ in range(20):
        eng.planes_all()
    out['all_terms_us_per_call'] = (time.perf_counter() - t0) / 20 * 1e6
    out['what'] = ('all_terms: arp_planes_run_all (one launch sequence: three cell grids, count, scan, emit; one wait) + the four '
                   'fetches, wall time per call; terms: each term run + fetched on its own')
    for name in ('ring_ring', 'atom_ring', 'amide_amide', 'amide_ring'):
        f = getattr(eng, name)
        n = int(f().shape[0])
        assert n == got[name].shape[0]
        t0 = time.perf_counter()
        for _ in range(10):
            f()
        entry = {'records': n, 'us_per_call': (time.perf_counter() - t0) / 10 * 1e6}
        if cpu:
            t0 = time.perf_counter()
            m = int(cpu_fn[name]().shape[0])
            entry['cpu_port_us'] = (time.perf_counter() - t0) * 1e6
            assert m == n
        out['terms'][name] = entry
    return out


def kdtree_leg(soa, cutoff):
    """a stronger CPU baseline for the SEARCH stage alone -- scipy's cKDTree (C++, float64) on the same
    coordinates, all unordered pairs within the cutoff, one core.  No filters, no classification."""
    try:
        from scipy.spatial import cKDTree
    except Exception:
        return None
    xyz = np.asarray(soa.xyz, dtype=np.float64)
    t0 = time.perf_counter()
    tree = cKDTree(xyz)
    pairs = tree.query_pairs(cutoff, output_type='ndarray')
    dt = time.perf_counter() - t0
    return {'pairs_within_cutoff': int(pairs.shape[0]), 'seconds': dt, 'pairs_per_s': pairs.shape[0] / dt, 'cores': 1,
            'what': 'scipy.spatial.cKDTree build + query_pairs on the same coordinates: neighbour search only (no filters, '
                    'no classification), the stage Bio.PDB.NeighborSearch.search_all performs in the reference'}


def dist_env():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    return rank, world, local


def run_reference(args):
    """The reference arm: the CPU port of the reference's loop on all host cores (rank 0 only)."""
    rank, world, _ = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.atoms                      # the same configs[2] cloud as the product arm, one per host thread
    port = CpuPort(sample, cores)
    for _ in range(max(args.warmup, 1)):
        port.step()
    t_tot, pairs = 0.0, 0
    for _ in range(args.steps):
        n, dt = port.step()
        pairs += n
        t_tot += dt
    value = pairs / t_tot
    desc = f'{cores} threads x one {sample}-atom cloud of the configs[2] recipe per step (C port of the reference loop; the reference itself is single-threaded Python)'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * t_tot / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32 distance / f64 angles / u32 masks', 'data': 'synthetic',
        'config': {'workload': workload_name(args.atoms), 'atoms_per_gpu': args.atoms, 'cutoff': 5.0, 'sample': desc},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    rank, world, local = dist_env()
    dist = None
    if world > 1:
        import torch.distributed as dist   # plumbing only: barrier + max/sum of per-rank scalars
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('gloo', rank=rank, world_size=world)

    from arpeggio_b200 import abi, params as arp_params, synth
    from arpeggio_b200.batch import BatchRunner
    from arpeggio_b200.engine import ContactEngine, PinnedBuffer, pinned_soa
    from arpeggio_b200.soa import AtomSoA

    p = arp_params.make_params()
    soa = synth.cloud_featured(args.atoms, seed=2 + rank)
    eng = ContactEngine(device=local, params=p)

    # ---- inputs resident in HBM ------------------------------------------------------------
    eng.upload_atoms(soa)
    n_pairs = eng.run_pairs()
    in_bytes = soa.input_bytes()
    alg_bytes = in_bytes + 16 * n_pairs

    sampler = ClockSampler(local) if rank == 0 else None
    eng.time_pairs(max(args.warmup, 3), flush_l2=True)
    if dist:
        dist.barrier()
    eng.sync()
    t_begin = time.time()
    ms_step = eng.time_pairs(args.steps, flush_l2=True)          # mean per-step CUDA-event time
    eng.sync()
    st = eng.stats()                                             # the split of the timed steps (before the next run resets it)
    l0 = eng.launch_count()
    eng.run_pairs()                                              # kernels of one step, counted by the library
    per_step = eng.launch_count() - l0
    launches = per_step * args.steps                             # the timed steps (the 32 split-timing steps not counted)
    if dist:
        dist.barrier()

    # ---- end to end through the public API, host buffers -----------------------------------
    # Every step moves its own inputs H2D from pinned host memory and its own results D2H.  Headline: the structure in
    # its wire form (soa.WireAtoms: uint8 counts for the CSR offsets, sparse halogen neighbours), the (i, j)-sorted
    # stream in its packed form (one 32-bit word per record + row offsets), pipelined over 8 streams from one host thread
    # (BatchRunner.run(packed=True): upload, kernels, sort, copies of a step enqueued without a host wait in between).
    pins = []
    host = pinned_soa(soa)                   # one pinned block: arp_upload_atoms moves it with a single DMA
    host_w = pinned_soa(soa.to_wire())
    out_pin = PinnedBuffer(16 * (n_pairs + 1024))
    out = out_pin.array(abi.PAIR_DTYPE)
    e2e_steps = min(max(10 * args.steps, 100), 600)
    runner = BatchRunner(device=local, slots=8, params=p)
    ref_rec = eng.fetch_pairs(n_pairs, sorted=True)                  # the 16-byte records of the resident run: the check of every leg

    def leg(src, steps, check=None, **kw):
        """warm-up (6 structures, results checked against the 16-byte records) + `steps` timed structures; seconds per step"""
        seen = []
        runner.run([src] * 8, consume=(lambda i, r: seen.append(check(r))) if check else None, check_finite=False, **kw)
        assert not check or (len(seen) == 8 and all(seen)), kw
        if dist:
            dist.barrier()
        _, dt = runner.run([src] * steps, check_finite=False, **kw)
        return dt / steps

    same = lambda rec: np.array_equal(rec[['i', 'j', 'mask']], ref_rec[['i', 'j', 'mask']])
    same_d = lambda rec: np.array_equal(rec, ref_rec)
    e2e_s = leg(host_w, e2e_steps, check=lambda r: same(r.to_records(soa.feat)), packed=True)
    in_bytes_wire = int(host_w.input_bytes())
    d2h_bytes = 4 * (soa.n_atoms + 2) + 4 * n_pairs + (n_pairs if soa.n_atoms > (1 << 17) else 0)      # row offsets + scratch entry, words
    legs = {}
    legs['plain_inputs'] = (leg(host, e2e_steps, check=lambda r: same(r.to_records(soa.feat)), packed=True), int(in_bytes), d2h_bytes,
                            'the AtomSoA as it is (int32 CSR offsets, dense neighbour array), packed stream')
    legs['with_distances'] = (leg(host_w, e2e_steps, check=lambda r: same_d(r.to_records(soa.feat)), packed=True, with_dist=True),
                              in_bytes_wire, d2h_bytes + 4 * n_pairs, 'wire inputs, packed stream + the float32 distance stream')
    steps_t = max(3, min(args.steps, 200))                           # the threaded legs ramp up in a few steps
    legs['compact'] = (leg(host, steps_t, check=lambda r: same(r.to_records()), compact=True), int(in_bytes), 4 * (soa.n_atoms + 1) + 8 * n_pairs,
                       'round-1/2 compact stream: 8-byte (j, mask) records + row offsets, one host thread per stream slot')
    legs['records16'] = (leg(host, steps_t, check=same_d, sorted=True), int(in_bytes), 16 * n_pairs,
                         '16-byte arp_pair records (BatchRunner.run(sorted=True))')
    src3 = synth.cloud_featured(args.atoms, seed=2 + rank, h_decimals=3)
    eng.upload_atoms(src3)
    ref3 = eng.fetch_pairs(eng.run_pairs(), sorted=True)
    host3 = pinned_soa(src3.to_wire())
    assert host3.h_fix is not None
    legs['wire_h_fix'] = (leg(host3, e2e_steps, check=lambda r: np.array_equal(r.to_records(src3.feat)[['i', 'j', 'mask']], ref3[['i', 'j', 'mask']]),
                              packed=True), int(host3.input_bytes()), 4 * (soa.n_atoms + 2) + 4 * ref3.shape[0],
                          'the same cloud with hydrogen coordinates of 3 decimals (PDB / mmCIF text precision): they travel as int32 '
                          'fixed point, checked lossless on the host')
    eng.upload_atoms(soa)
    assert eng.run_pairs() == n_pairs
    # one stream, one structure at a time (latency, not throughput)
    sbuf = runner._packed_buffer(0, soa.n_atoms, n_pairs + 1024, False)

    def e2e_serial():
        eng.upload_atoms(host_w, check_finite=False)
        eng.run_pairs_async()
        eng.fetch_pairs_packed_async(sbuf, n_pairs, False)
        return eng.fetch_pairs_packed_wait()

    for _ in range(3):
        assert e2e_serial().n == n_pairs
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(steps_t):
        e2e_serial()
    e2e_serial_s = (time.perf_counter() - t0) / steps_t
    got = ref_rec
    # resident inputs, steps of THREE contexts (streams) in flight at once: what the device sustains when the launch ramps and
    # tails of consecutive steps overlap (the per-step figure `value` serialises them); every context holds its own copy of
    # the structure and of all buffers (3 x 27 MB in flight), rank 0 only
    pipelined = None
    if rank == 0:
        engs = [ContactEngine(device=local, params=p) for _ in range(3)]
        for e in engs:
            e.upload_atoms(soa)
            assert e.run_pairs() == n_pairs
        reps = max(args.steps, 30)

        def spin(e):
            for _ in range(reps):
                e.run_pairs_async()
            e.sync()

        for e in engs:
            spin(e)
        ths = [threading.Thread(target=spin, args=(e,)) for e in engs]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        dtp = time.perf_counter() - t0
        pipelined = {'value': 3 * reps * n_pairs / dtp, 'unit': UNIT, 'us_per_step': dtp / (3 * reps) * 1e6, 'contexts': 3,
                     'what': 'inputs resident, three contexts on three streams run the step concurrently, wall clock over all steps, no L2 '
                             'flush (the three working sets are 81 MB): throughput with the ramps and tails of consecutive steps overlapped'}
        for e in engs:
            e.close()
    # what PCIe gives this rank while all ranks copy at once: the same sizes, H2D and D2H concurrently
    if dist:
        dist.barrier()
    pcie = eng.memcpy_probe(in_bytes_wire, d2h_bytes, 40)
    # ---- configs[4]: PDB-batch throughput, this rank's shard of 20k-atom structures, host buffers, end to end ----
    batch = None
    if args.batch_structures > 0:
        distinct = [pinned_soa(synth.cloud_featured(args.batch_atoms, seed=1000 + 97 * rank + k).to_wire()) for k in range(8)]
        shard = [distinct[k % len(distinct)] for k in range(args.batch_structures)]
        # warm-up: two launch sequences per stream slot (the first sizes the record buffers, the second the sorted views from them)
        runner.run((shard * 2)[:2 * len(runner.engines) * args.batch_pack], check_finite=False, packed=True, pack=args.batch_pack)
        if dist:
            dist.barrier()
        counts, dt_b = runner.run(shard, check_finite=False, packed=True, pack=args.batch_pack)
        batch = (len(shard), float(sum(counts)), dt_b)
    runner.close()
    large = None
    if rank == 0 and args.large_atoms > 0:
        big = synth.cloud_featured(args.large_atoms, seed=5)
        eng.upload_atoms(big)
        n_big = eng.run_pairs()
        eng.time_pairs(3, flush_l2=True)
        ms_big = eng.time_pairs(20, flush_l2=True)
        large = (args.large_atoms, n_big, ms_big, big.input_bytes() + 16 * n_big)
        del big
        eng.upload_atoms(soa)
        eng.run_pairs()
    json_leg = json_emitter_leg(got, args.atoms) if rank == 0 and not args.no_cpu else None
    planes_leg = planes_leg_run(eng, soa, p, args.atoms, cpu=not args.no_cpu) if rank == 0 and args.atoms >= 1000 else None
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if sampler else None

    # ---- aggregate over ranks: slowest rank's time, total pairs ------------------------------
    ms_pair = st['ms_pairs']                # the three pair kernels back to back, one event before and one after
    tot_pairs, ms_max, e2e_max, e2e_serial_max = float(n_pairs), ms_step, e2e_s, e2e_serial_s
    leg_names = sorted(legs)
    leg_max = {k: legs[k][0] for k in leg_names}
    pcie_min = dict(pcie)
    if dist:
        import torch
        t = torch.tensor([ms_step, e2e_s, e2e_serial_s, batch[2] if batch else 0.0, -pcie['h2d_gbs'], -pcie['d2h_gbs']] + [legs[k][0] for k in leg_names],
                         dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pcie_min = {'h2d_gbs': -float(t[4]), 'd2h_gbs': -float(t[5])}
        leg_max = {k: float(t[6 + n]) for n, k in enumerate(leg_names)}
        s = torch.tensor([float(n_pairs), batch[0] if batch else 0.0, batch[1] if batch else 0.0], dtype=torch.float64)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        ms_max, e2e_max, e2e_serial_max, tot_pairs = float(t[0]), float(t[1]), float(t[2]), float(s[0])
        if batch:
            batch = (int(s[1]), float(s[2]), float(t[3]))

    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = alg_bytes / (ms_max * 1e-3) / 1e9            # t = first kernel start to last kernel end
        achieved_pairs = alg_bytes / (ms_pair * 1e-3) / 1e9
        line = {
            'metric': METRIC, 'value': tot_pairs / (ms_max * 1e-3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms_max, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32 distance / f64 angles / u32 masks', 'data': 'synthetic',
            'config': {'workload': workload_name(args.atoms), 'atoms_per_gpu': args.atoms, 'pairs_per_structure': n_pairs,
                       'cutoff': 5.0, 'l2': 'flushed between steps (384 MiB memset outside the event brackets)',
                       'timing': 'sum of per-step CUDA-event brackets on the library stream, max over ranks',
                       'sharding': 'one independent structure per GPU, no collective'},
            'e2e': {'value': tot_pairs / e2e_max, 'unit': UNIT, 'h2d_bytes_per_step': in_bytes_wire,
                    'd2h_bytes_per_step': d2h_bytes, 'steps': e2e_steps, 'ms_per_step': e2e_max * 1e3, 'sorted': True,
                    'inputs': 'soa.WireAtoms in one pinned block: uint8 per-atom counts in place of the two int32 CSR offset arrays, the '
                              'halogen neighbours as (index, coordinate) rows; decoded on the device after the copy',
                    'stream': 'packed: uint32 row offsets [atoms + 1] + one 32-bit word per record (j | SIFt bits << 17), (i, j) ascending; '
                              'the entity class is recomputed on the host from the feat words, float32 distances stay on the device '
                              '(arp_pairs_fetch_dist)',
                    'api': 'BatchRunner.run(packed=True): 8 contexts (streams), one host thread enqueues every step whole (upload, kernels, '
                           'sort, copies: arp_pairs_run_async + arp_pairs_fetch_packed_async) and waits for it when its slot comes round again',
                    'serial_value': tot_pairs / e2e_serial_max, 'serial_ms_per_step': e2e_serial_max * 1e3,
                    'serial_api': 'one context: upload_atoms + run_pairs_async + fetch_pairs_packed_async + fetch_pairs_packed_wait per structure',
                    'legs': {k: {'value': tot_pairs / leg_max[k], 'ms_per_step': leg_max[k] * 1e3, 'h2d_bytes_per_step': legs[k][1],
                                 'd2h_bytes_per_step': int(legs[k][2]), 'what': legs[k][3]} for k in leg_names},
                    'pcie': dict(pcie_min, what='pinned cudaMemcpyAsync of the step\'s H2D and D2H sizes, both directions at once, '
                                                'every rank at the same time (min over ranks, GB/s per GPU)',
                                 d2h_floor_ms=d2h_bytes / (pcie_min['d2h_gbs'] * 1e6) if pcie_min['d2h_gbs'] else None,
                                 h2d_floor_ms=in_bytes_wire / (pcie_min['h2d_gbs'] * 1e6) if pcie_min['h2d_gbs'] else None,
                                 e2e_fraction_of_floor=(max(d2h_bytes / pcie_min['d2h_gbs'], in_bytes_wire / pcie_min['h2d_gbs']) / 1e6 / (e2e_max * 1e3)
                                                        if pcie_min['d2h_gbs'] and pcie_min['h2d_gbs'] else None))},
            'resident_pipelined': pipelined,
            'gpu_launches': int(launches),
            'kernels_per_step': int(per_step),
            'roofline': {'bound': 'hbm', 'kernel': 'the whole step: k_grid_reg + k_search + k_classify + k_hscan',
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         'frac_of_nominal_8tbs': achieved / 8000.0,
                         'traffic': ncu_traffic(args.atoms)[0], 'traffic_l2_bytes': ncu_traffic(args.atoms)[1],
                         'traffic_source': 'profiles/k_pairs_traffic.json: ncu captures of this build (tools/prof.sh); DRAM bytes with caches flushed per kernel, L2 bytes with caches left alone',
                         'algorithmic_bytes': int(alg_bytes), 'kernel_ms': ms_max,
                         'pair_kernels': {'kernel': 'k_search + k_classify + k_hscan, one event before and one after',
                                          'kernel_ms': ms_pair, 'achieved': achieved_pairs, 'frac': achieved_pairs / peak},
                         'grid_build_ms': st['ms_grid'],
                         'split_with_events_between_all_kernels': {'search_ms': st['ms_search'], 'classify_ms': st['ms_classify'] - st['ms_hscan'], 'hscan_ms': st['ms_hscan']},
                         'peak_source': peak_src},
            'clocks': clocks,
            'candidate_tests_per_step': int(st['n_candidates']),
            'candidate_tests_per_s': float(st['n_candidates']) * world / (ms_max * 1e-3),
        }
        if large:
            a_l = large[3] / (large[2] * 1e-3) / 1e9
            line['roofline']['large'] = {'atoms': large[0], 'pairs': large[1], 'ms_per_step': large[2], 'pairs_per_s': large[1] / (large[2] * 1e-3),
                                         'algorithmic_bytes': int(large[3]), 'achieved': a_l, 'frac': a_l / peak,
                                         'what': 'the same step on a 1M-atom cloud of the same recipe (20 steps, L2 flushed), where launch ramps and tails no longer dominate'}
        if json_leg:
            line['json'] = json_leg
        if planes_leg:
            line['planes'] = planes_leg
        if batch:
            line['batch'] = {'metric': 'structures/s (configs[4]: PDB-batch of synthetic 20k-atom structures)',
                             'value': batch[0] / batch[2], 'unit': 'structures/s', 'structures': batch[0],
                             'atoms_per_structure': args.batch_atoms, 'pairs_per_s': batch[1] / batch[2],
                             'seconds': batch[2], 'structures_per_launch': args.batch_pack,
                             'sharding': f'{args.batch_structures} structures per GPU, {args.batch_pack} per launch sequence (one DMA per '
                                         'structure in wire form, concatenated on the device: arp_upload_atoms_batch), 8 stream slots driven by one host thread, '
                                         'no collective; H2D + kernels + sort + packed D2H (5 bytes per record) of every structure inside the timed region'}
        if not args.no_cpu:
            kd = kdtree_leg(soa, 5.0)
            if kd:
                line['cpu_kdtree_search_only'] = kd
            cores = os.cpu_count() or 1
            port = CpuPort(args.atoms, cores)
            n_cpu, dt = port.step(3)
            line['cpu_baseline'] = {'value': n_cpu / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                    'sample': f'{cores} threads x 3 passes over one {args.atoms}-atom cloud of the same recipe each '
                                              f'({cores * dt:.0f} s of CPU time, {dt:.1f} s wall); C port of the reference loop'}
        print(json.dumps(line), flush=True)
    eng.close()
    for pb in pins:
        pb.free()
    out_pin.free()
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=300)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', choices=('ours', 'reference'), default='ours')
    ap.add_argument('--atoms', type=int, default=100_000)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--batch-structures', type=int, default=768, help='structures per GPU of the PDB-batch leg (0: skip)')
    ap.add_argument('--batch-atoms', type=int, default=20_000)
    ap.add_argument('--batch-pack', type=int, default=16, help='structures per launch sequence of the PDB-batch leg')
    ap.add_argument('--large-atoms', type=int, default=1_000_000, help='size of the roofline.large leg (0: skip)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
