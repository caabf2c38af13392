import os, sys, json, torch
sys.path.insert(0,'/root/repo')
from tools.nas_resident_check import build, timing, parity
for env in ({}, {"HN_NAS_DW_F32":"1"}):
    for arch in ("wang2","wang3","wang4"):
        for k in ("HN_NAS_DW_F32",): os.environ.pop(k,None)
        os.environ.update(env)
        r=timing(arch, {}); print("TIMING", arch, env, "%.3f ms %.2f M/s"%(r["ms"], r["patches_per_sec"]/1e6))
