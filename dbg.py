import numpy as np, torch, sys
sys.path.insert(0,'.')
import neon_b200 as nb
from neon_b200 import problems as P
bk = nb.Backend()
grid = nb.dGrid(bk, (64, 8, 8))
pop0, pop1, flag = P.setup_device(grid, 19, np.float32, 0)
c = nb.LbmContainers.iteration(nb.StencilSemantic.streaming, pop0, pop1, flag, 1.3, opts=nb.opt_kernel(nb.KERNEL_TMA))
c.run(0, nb.DataView.STANDARD)
bk.syncAll()
print("ok", float(pop1.data.sum()))
