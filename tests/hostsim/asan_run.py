"""TEST INFRASTRUCTURE: every solver path (ADMM + polish, warm polish, interior point, robust chain, quadruped, large and small
trees) of the single-lane host build under AddressSanitizer.  Run by tests/test_hostsim_asan.py in a subprocess with
libasan preloaded; argv[1] = the instrumented library."""
import sys, numpy as np, ctypes as C
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "belief-planning_b200"))
from _bmpc import scenarios, abi
from tests.hostsim import driver
from tests.helpers import force_interior_point
driver._lib = C.CDLL(sys.argv[1]); driver._lib.hostsim_last_error.restype = C.c_char_p
B=12
x0,z0,xref,pp=scenarios.highway_batch(B,seed=5)
for forced in (False, True):
    for ctrl in (abi.CTRL_BRANCH, abi.CTRL_ROBUST):
        cfg=scenarios.highway_config(); cfg.controller=ctrl
        if forced: force_interior_point(cfg)
        hs=driver.HostSim(cfg,B)
        for s in range(2):
            r=hs.solve(x0,z0,xref,pp)
        print("highway ctrl",ctrl,"forced",forced,np.bincount(r["status"],minlength=4))
    q0,qz,qr=scenarios.quadruped_batch(8,seed=2)
    cfg=scenarios.quadruped_config()
    if forced: force_interior_point(cfg)
    hs=driver.HostSim(cfg,8); r=hs.solve(q0,qz,qr); r=hs.solve(q0,qz,qr); print("quad forced",forced,np.bincount(r["status"],minlength=4))
for m,NB in ((4,3),(2,1)):
    names=["maintain","brake","lc","trackv"][:m]
    cfg=force_interior_point(scenarios.highway_config(policies=names,NB=NB))
    p=np.zeros((4,m,4)); 
    if m>=3: p[:,2,:]=pp[:4,2,:]
    if m>=4: p[:,3,0]=20
    hs=driver.HostSim(cfg,4); r=hs.solve(x0[:4],z0[:4],xref[:4],p); print("sweep",m,NB,np.bincount(r["status"],minlength=4))
print("ASAN_RUN_OK")
