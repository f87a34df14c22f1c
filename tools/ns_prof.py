import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch
from bench import make_inputs, ns_iteration_seconds, C5_BOX, _EfpOwner
from pyds_b200 import capi, models
from pyds_b200.parallel import ShardedEfp
eng6 = capi.Engine(models.kinetics("design6").lower().blob, 0)
sh6 = ShardedEfp(eng6, 1, 0)
_, th5, w5 = make_inputs(10, 100000)
for dev in ("cuda:0", None, "cuda:0"):
    ns_iteration_seconds(sh6, C5_BOX, 2000, th5[:2000], np.full(2000, 1/2000), 2, dev)
    t=time.time(); r = ns_iteration_seconds(sh6, C5_BOX, 100000, th5, w5, 3, dev); print(dev, r, time.time()-t, flush=True)
