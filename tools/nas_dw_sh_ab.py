import os, sys
sys.path.insert(0,'/root/repo')
from tools.nas_resident_check import timing
for sh8 in ("1","0"):
    os.environ["HN_NAS_DW_SH8"]=sh8
    for arch in ("wang2","wang3","wang4"):
        r=timing(arch,{}); print("DWSH", arch, "sh8="+sh8, "%.3f ms %.2f M/s"%(r["ms"], r["patches_per_sec"]/1e6), flush=True)
