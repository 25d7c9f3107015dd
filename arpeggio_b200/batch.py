"""Whole-batch runs: independent structures, one per stream slot, sharded over the GPUs of a box.

The reference processes one structure per process invocation (process_protein_cli.py:153-198);
structures share no state, so a batch shards with no collective on the data path:
  - across GPUs: `shard_indices` gives every rank (one process per GPU) its part of the list,
    greedy longest-first by atom count so that the ranks finish together;
  - inside a GPU: `BatchRunner` keeps `slots` contexts (device buffers + stream each) busy, so the
    H2D copy of one structure, the kernels of another and the D2H copy of a third overlap (PCIe is
    full duplex).  The packed stream (`run(packed=True)`) is pipelined from ONE host thread: a step is
    enqueued whole (upload, kernels, sorted view, copies) and waited for when its slot comes round
    again -- several submitting threads only contend for the driver's launch path.  The older stream
    forms (16-byte records, compact) use one worker thread per slot (ctypes releases the GIL).
"""
import queue
import threading
import time

import numpy as np

from . import abi
from .engine import CompactPairs, ContactEngine, PackedPairs, PinnedBuffer


def shard_indices(sizes, world_size, rank):
    """Indices of the structures rank `rank` of `world_size` processes.  sizes: atoms per structure.
    Greedy longest-processing-time assignment (deterministic: ties by index), returned ascending."""
    if not 0 <= rank < world_size:
        raise ValueError('rank out of range')
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.lexsort((np.arange(sizes.shape[0]), -sizes))
    load = [0] * world_size
    mine = []
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += int(sizes[i]) + 1
        if r == rank:
            mine.append(int(i))
    return sorted(mine)


class BatchRunner:
    """`slots` ContactEngines (contexts, one stream each) on one device.  run(packed=True) drives them from the calling
    thread (submit_threads of them when asked for), every step enqueued whole; the other stream forms use one worker
    thread per slot."""

    def __init__(self, device=0, slots=8, params=None, submit_threads=1):
        self.device = device
        self.submit_threads = submit_threads          # host threads that enqueue the pipelined packed stream (run(packed=True))
        self.engines = [ContactEngine(device, params) for _ in range(max(1, slots))]
        self._pins = [None] * len(self.engines)
        self._last_n = [0] * len(self.engines)
        self._ratio = [0.0] * len(self.engines)       # records per atom of the slot's last structure (sizes the blind copy)

    def close(self):
        for e in self.engines:
            e.close()
        for p in self._pins:
            if p is not None:
                p.free()
        self.engines, self._pins = [], []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_params(self, p):
        for e in self.engines:
            e.set_params(p)

    def _out_buffer(self, slot, n):
        pb = self._pins[slot]
        if pb is None or pb.nbytes < 16 * n:
            if pb is not None:
                pb.free()
            pb = self._pins[slot] = PinnedBuffer(16 * (n + n // 8 + 1024))
        return pb.array(abi.PAIR_DTYPE)

    def _compact_buffer(self, slot, n_atoms, n, with_dist):
        """One pinned block per slot carved into row offsets | compact records | distances."""
        o_rec = (4 * (n_atoms + 1) + 255) // 256 * 256
        o_dist = o_rec + (8 * n + 255) // 256 * 256
        need = o_dist + (4 * n if with_dist else 0)
        pb = self._pins[slot]
        if pb is None or pb.nbytes < need:
            if pb is not None:
                pb.free()
            pb = self._pins[slot] = PinnedBuffer(need + need // 8 + 4096)
        raw = pb.array(np.uint8)
        cap = (pb.nbytes - o_rec) // (12 if with_dist else 8)
        o_dist = o_rec + (8 * cap + 255) // 256 * 256
        if with_dist:
            cap = min(cap, (pb.nbytes - o_dist) // 4)
        return CompactPairs(raw[:4 * (n_atoms + 1)].view(np.uint32), raw[o_rec:o_rec + 8 * cap].view(abi.PAIR_C_DTYPE),
                            raw[o_dist:o_dist + 4 * cap].view(np.float32) if with_dist else None)

    def _packed_buffer(self, slot, n_atoms, n, with_dist, largest=None):
        """One pinned block per slot carved into row offsets | low words | high bytes | distances.  largest: atoms of the
        largest structure when n_atoms is the total of a batch (the words hold structure-local indices)."""
        wide = (n_atoms if largest is None else largest) > (1 << 17)
        per = 4 + (1 if wide else 0) + (4 if with_dist else 0)
        o_lo = (4 * (n_atoms + 2) + 255) // 256 * 256          # + 1: scratch entry of arp_pairs_fetch_packed_async
        need = o_lo + per * n + 1024
        pb = self._pins[slot]
        if pb is None or pb.nbytes < need:
            if pb is not None:
                pb.free()
            pb = self._pins[slot] = PinnedBuffer(need + need // 8 + 4096)
        raw = pb.array(np.uint8)
        cap = (pb.nbytes - o_lo - 1024) // per
        o_hi = (o_lo + 4 * cap + 255) // 256 * 256
        o_dist = (o_hi + (cap if wide else 0) + 255) // 256 * 256
        return PackedPairs(raw[:4 * (n_atoms + 2)].view(np.uint32), raw[o_lo:o_lo + 4 * cap].view(np.uint32),
                           raw[o_hi:o_hi + cap] if wide else None,
                           raw[o_dist:o_dist + 4 * cap].view(np.float32) if with_dist else None)

    def _run_packed(self, soas, consume, with_dist, check_finite, counts, pack=1, slots=None, units=None):
        """The packed stream, pipelined from ONE host thread: every structure (or group of `pack` structures, uploaded with
        arp_upload_atoms_batch) is enqueued whole -- upload, kernels, sorted packed view, copies -- on the next slot's
        stream without a host wait in between (arp_pairs_fetch_packed_async), and waited for only when its slot comes
        round again.  Six threads doing the same contend for the driver's launch path (each enqueue call then takes 5-10x
        longer, profiles/README.md round 2); one thread enqueues a whole step in well under 100 us.  slots / units: the
        share of one submission thread when there are several (submit_threads)."""
        slots = list(range(len(self.engines))) if slots is None else slots
        units = range(0, len(soas), pack) if units is None else units
        S = len(slots)
        pending = {s: None for s in slots}

        def finish(slot):
            i, off, largest = pending[slot]
            pending[slot] = None
            n_atoms = int(off[-1])
            rec = self.engines[slot].fetch_pairs_packed_wait(grow=lambda m, s=slot, a=n_atoms: self._packed_buffer(s, a, m, with_dist, largest))
            if n_atoms:
                self._ratio[slot] = rec.n / n_atoms
            for k in range(len(off) - 1):
                part = rec if len(off) == 2 else rec.structure(int(off[k]), int(off[k + 1]))
                counts[i + k] = part.n
                if consume is not None:
                    consume(i + k, part)

        k = -1
        for k, i in enumerate(units):
            slot = slots[k % S]
            if pending[slot] is not None:
                finish(slot)
            eng = self.engines[slot]
            if pack > 1:
                off = eng.upload_atoms_batch(soas[i:i + pack], check_finite=check_finite)
            else:
                eng.upload_atoms(soas[i], check_finite=check_finite)
                off = (0, soas[i].n_atoms)
            n_atoms = int(off[-1])
            eng.run_pairs_async()
            expect = int(self._ratio[slot] * n_atoms * 1.02) + 64 if self._ratio[slot] else 14 * n_atoms
            largest = eng.max_struct_atoms()
            buf = self._packed_buffer(slot, n_atoms, max(expect, 14 * n_atoms), with_dist, largest)
            eng.fetch_pairs_packed_async(buf, expect, with_dist)
            pending[slot] = (i, off, largest)
        for j in range(S):                                  # oldest first
            slot = slots[(k + 1 + j) % S]
            if pending[slot] is not None:
                finish(slot)

    def run(self, soas, consume=None, sorted=False, check_finite=True, compact=False, with_dist=False, pack=1, packed=False):
        """Upload -> grid build + pair kernels -> fetch for every AtomSoA of `soas`.

        consume(index, records): called in the worker thread with a view of the slot's pinned record
        buffer (valid only during the call).  compact: the (i, j)-sorted stream in its compact form
        (engine.CompactPairs; distances only with_dist) instead of 16-byte records.  pack > 1 (compact only): that
        many consecutive structures go up as ONE batch (arp_upload_atoms_batch: one DMA per structure, concatenated on
        the device) and run as one launch sequence; consume still sees one CompactPairs per structure (the j of its
        records are batch-global: subtract its atom_base).  packed: the stream as engine.PackedPairs, 4 or 5 bytes per
        record, pipelined from the calling thread (consume runs there too; with pack > 1 it sees one PackedPairs per
        structure whose to_records gives structure-local indices).
        Returns (pairs_per_structure, seconds)."""
        soas = list(soas)
        counts = [0] * len(soas)
        todo = queue.SimpleQueue()
        pack = max(1, int(pack)) if (compact or packed) else 1
        for i in range(0, len(soas), pack):
            todo.put(i)
        errors = []

        def work(slot):
            eng = self.engines[slot]
            try:
                while True:
                    try:
                        i = todo.get_nowait()
                    except queue.Empty:
                        return
                    if pack > 1:
                        group = soas[i:i + pack]
                        off = eng.upload_atoms_batch(group, check_finite=check_finite)
                        eng.run_pairs_async()
                        n_atoms = int(off[-1])
                        guess = max(self._last_n[slot], 14 * n_atoms)
                        buf = self._compact_buffer(slot, n_atoms, guess, with_dist)
                        rec = eng.fetch_pairs_compact(with_dist, out=buf, grow=lambda m, s=slot, a=n_atoms:
                                                      self._compact_buffer(s, a, m, with_dist))
                        self._last_n[slot] = rec.n
                        for k in range(len(group)):
                            part = rec.structure(int(off[k]), int(off[k + 1]))
                            counts[i + k] = part.n
                            if consume is not None:
                                consume(i + k, part)
                        continue
                    eng.upload_atoms(soas[i], check_finite=check_finite)
                    if compact:
                        # one wait per structure: the fetch waits for the run; the slot's buffer is sized from the
                        # structures it has seen (a stream that does not fit reports its length and is fetched again)
                        eng.run_pairs_async()
                        guess = max(self._last_n[slot], 14 * soas[i].n_atoms)
                        buf = self._compact_buffer(slot, soas[i].n_atoms, guess, with_dist)
                        rec = eng.fetch_pairs_compact(with_dist, out=buf, grow=lambda m, s=slot, a=soas[i].n_atoms:
                                                      self._compact_buffer(s, a, m, with_dist))
                        n = self._last_n[slot] = rec.n
                    else:
                        n = eng.run_pairs()
                        rec = eng.fetch_pairs(n, sorted=sorted, out=self._out_buffer(slot, n))
                    counts[i] = n
                    if consume is not None:
                        consume(i, rec)
            except Exception as err:           # surface the first failure in the caller's thread
                errors.append(err)

        if packed:
            t0 = time.perf_counter()
            T = max(1, min(self.submit_threads, len(self.engines)))
            if T == 1:
                self._run_packed(soas, consume, with_dist, check_finite, counts, pack)
            else:
                def share(t):
                    try:
                        self._run_packed(soas, consume, with_dist, check_finite, counts, pack, list(range(t, len(self.engines), T)),
                                         range(t * pack, len(soas), T * pack))
                    except Exception as err:
                        errors.append(err)
                ths = [threading.Thread(target=share, args=(t,)) for t in range(T)]
                for t in ths:
                    t.start()
                for t in ths:
                    t.join()
                if errors:
                    raise errors[0]
            for e in self.engines:
                e.sync()
            return counts, time.perf_counter() - t0
        t0 = time.perf_counter()
        threads = [threading.Thread(target=work, args=(s,)) for s in range(len(self.engines))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in self.engines:
            e.sync()
        dt = time.perf_counter() - t0
        if errors:
            raise errors[0]
        return counts, dt
