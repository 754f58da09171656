"""Fingerprints of the lowered tables of the GPU-verified scenarios: `python tools/hash_tables.py <repo root>`
for two checkouts shows whether a host-side change altered what the kernels consume (used to confirm that the
work done without a GPU left the tables of BASELINE configs 1-5 bit-identical)."""
import sys, hashlib
import numpy as np
sys.path.insert(0, sys.argv[1])
from omg_tools_b200 import scenarios as sc
for name in ('config1','config2','config4','config5','config_dubins','config_freeT','config_holonomic3d','config_quadrotor2d'):
    tb = getattr(sc,name)(build_solver=False).father.tables
    h = hashlib.md5()
    for tl in (tb.G, tb.J, tb.W, tb.F, tb.DF):
        for a in (tl.ptr, tl.coef, tl.cidx, tl.xi, tl.lrow): h.update(np.ascontiguousarray(a).tobytes())
    for a in (tb.tape_coef, tb.tape_fac, tb.tape_func, tb.kkt_pos_var, tb.env_ptr, tb.hp_s1, tb.hp_s2, tb.lbg, tb.ubg):
        h.update(np.ascontiguousarray(a).tobytes())
    print(name, h.hexdigest()[:16])
cfg3 = sc.config3(4, build_solver=False).tb
h = hashlib.md5()
for tl in (cfg3.G, cfg3.J, cfg3.W): 
    for a in (tl.ptr, tl.coef, tl.cidx, tl.xi): h.update(np.ascontiguousarray(a).tobytes())
print('config3', h.hexdigest()[:16])
