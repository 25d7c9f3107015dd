import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from arpeggio_b200 import params, synth
from arpeggio_b200.engine import ContactEngine
# a protein-like bond table: every atom bonded to its two sequence neighbours inside its residue (8 atoms per residue)
soa = synth.cloud_featured(100000, seed=2)
n = soa.n_atoms
nbr = [[] for _ in range(n)]
for i in range(n - 1):
    if soa.res_id[i] == soa.res_id[i + 1]:
        nbr[i].append(i + 1); nbr[i + 1].append(i)
off = np.zeros(n + 1, np.int32); off[1:] = np.cumsum([len(x) for x in nbr])
soa.bond_off = off; soa.bond_nbr = np.array([j for x in nbr for j in x], np.int32)
with ContactEngine(0, params.make_params()) as eng:
    eng.upload_atoms(soa); npairs = eng.run_pairs()
    eng.time_pairs(20, flush_l2=True)
    ms = eng.time_pairs(200, flush_l2=True); st = eng.stats()
    print('%-6s bonds on every atom: us/step %6.1f | classify alone %5.1f | %d records' % (sys.argv[1], ms * 1e3, (st['ms_classify'] - st['ms_hscan']) * 1e3, npairs))
