import time, sys
sys.path.insert(0,'/root/repo')
from amclj_b200 import _lib, maps
g = maps.lgfloor()
c = _lib.Context(0, "NS")
c.set_map(g); c.synchronize()
for _ in range(3):
    t=time.perf_counter(); c.set_map(g); c.synchronize(); print('set_map (auto: isotropic + 32 directional planes) ms', 1e3*(time.perf_counter()-t))
c.set_params(_lib.params("NS", df_mode=1))
t=time.perf_counter(); c.set_map(g); c.synchronize(); print('set_map (isotropic only) ms', 1e3*(time.perf_counter()-t))
