"""diagnostic: racecheck over the launch chain / the resident grid separately (tools/r02q.sh)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neon_b200 as nb
from neon_b200 import problems as P

mode = sys.argv[1]
bk = nb.Backend()
grid = nb.dGrid(bk, (44, 18, 12))
pop0, pop1, flag = P.setup_device(grid, 19, np.float32, P.CAVITY_SPHERE)
opts = {"chain": 0, "chain_all": 15 << 16, "step": 0, "chain_nospec": (1 << 29), "chain_nokeep": (1 << 27) | (1 << 29)}[mode]  # ("coop": the retired resident-grid kernel, profiles/r02q_race.log)
it = nb.LbmIteration(nb.StencilSemantic.streaming, nb.Occ.none, nb.TransferMode.get, pop0, pop1, flag, 1.3, lattice_q=19, arith=nb.ARITH_FAST, opts=opts)
if mode == "step":
    for _ in range(3):
        it.run()
else:
    it.runMany(3)
bk.syncAll()
print("probe", mode, "done", np.isfinite(it.getInput().updateHostData()).all())
