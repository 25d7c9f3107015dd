"""Contact JSON without one Python object per contact.

The reference builds a dict per contact in ``InteractionComplex.get_contacts`` (interactions.py:172-212)
and writes the list with ``json.dump(contacts, fp, indent=4, sort_keys=True)``
(process_protein_cli.py:187-188): about 40 us of interpreter time per contact.  Here every ATOM is
rendered once by the host's own ``utils.make_pymol_json`` and ``json.dumps`` (so escaping and key order
are Python's), and the C emitter of libarpeggio_cuda.so (``arp_pairs_json_write``, csrc/arp_json.cu)
joins the pieces for all atom-atom records on several host threads; the few plane / group entries go
through the reference's own code.  The text is byte-identical to the reference's dump.
"""
import ctypes as C
import functools
import json
import os

import numpy as np

from . import abi
from ._lib import ArpeggioCudaError, lib


@functools.lru_cache(maxsize=1 << 16)
def _scalar_text(kind, value):
    return json.dumps(value)


def atom_fragment(atom_dict):
    """The 'bgn' / 'end' object of one atom as json.dump(indent=4, sort_keys=True) writes it at depth 2."""
    if not atom_dict:
        return b'{}'
    parts = []
    for k in sorted(atom_dict):
        v = atom_dict[k]
        if type(v) not in (str, int, bool, float, type(None)) or type(k) is not str:
            # anything but flat scalars: let the json module lay it out, then shift it to depth 2
            lines = json.dumps(atom_dict, indent=4, sort_keys=True).split('\n')
            return '\n'.join([lines[0]] + ['        ' + ln for ln in lines[1:]]).encode('ascii')
        parts.append('            ' + _scalar_text(str, k) + ': ' + _scalar_text(type(v), v))
    return ('{\n' + ',\n'.join(parts) + '\n        }').encode('ascii')


def pairs_json(records, fragments, threads=None):
    """The atom-atom entries of the dump for `records` (arp_pair array, host), joined by ",\\n", without the
    enclosing brackets, as a uint8 NumPy buffer (bytes-like; not zero-filled first, the writer threads touch
    the pages).  fragments[a] = atom_fragment(...) of atom a."""
    rec = np.ascontiguousarray(records, dtype=abi.PAIR_DTYPE)
    n, n_atoms = rec.shape[0], len(fragments)
    if n == 0:
        return np.zeros(0, dtype=np.uint8)
    threads = int(threads or min(16, os.cpu_count() or 1))
    L = lib()
    lens = np.fromiter(map(len, fragments), dtype=np.uint32, count=n_atoms)
    ptrs = (C.c_char_p * n_atoms)(*fragments)
    size = C.c_uint64()
    rc = L.arp_pairs_json_size(rec.ctypes.data, n, n_atoms, lens.ctypes.data, threads, C.byref(size))
    if rc != abi.OK:
        raise ArpeggioCudaError(rc, 'arp_pairs_json_size: a record refers to an atom outside the fragment list')
    out = np.empty(size.value, dtype=np.uint8)
    written = C.c_uint64()
    rc = L.arp_pairs_json_write(rec.ctypes.data, n, n_atoms, C.cast(ptrs, C.c_void_p), lens.ctypes.data, threads,
                                out.ctypes.data, size.value, C.byref(written))
    if rc != abi.OK or written.value != size.value:
        raise ArpeggioCudaError(rc, 'arp_pairs_json_write failed')
    return out


def _rest_text(other_entries):
    rest = json.dumps(other_entries, indent=4, sort_keys=True) if other_entries else ''
    return rest[2:-2] if rest else ''                   # strip "[\n" and "\n]"


def write_spliced(fp, atom_atom, other_entries):
    """Write the whole dump to the binary file `fp` without turning the atom-atom text into a str."""
    rest = _rest_text(other_entries).encode('ascii')
    if not len(atom_atom) and not rest:
        fp.write(b'[]')
        return
    fp.write(b'[\n')
    if len(atom_atom):
        fp.write(memoryview(atom_atom))
        if rest:
            fp.write(b',\n')
    fp.write(rest)
    fp.write(b'\n]')


def splice(atom_atom_bytes, other_entries):
    """The whole dump: atom-atom entries (already text) followed by `other_entries` (list of dicts)."""
    rest = _rest_text(other_entries)
    if not len(atom_atom_bytes) and not rest:
        return '[]'
    body = bytes(atom_atom_bytes).decode('ascii') if len(atom_atom_bytes) else ''
    if body and rest:
        body += ',\n'
    return '[\n' + body + rest + '\n]'
