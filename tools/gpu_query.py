import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from btb_phylo_b200 import _capi
import torch
torch.cuda.init()
lib = _capi.lib()
for k in (b"umma_max_active_clusters_cg2", b"umma_max_active_clusters_cg1", b"umma_cta_group"):
    v = ctypes.c_int64(-1)
    rc = lib.btb_query(k, ctypes.byref(v))
    print(k.decode(), rc, v.value, _capi.last_error() if rc else "")
p = torch.cuda.get_device_properties(0)
print(p.name, p.multi_processor_count, p.total_memory)
print("cpus", os.cpu_count())
os.system("free -g | head -2; df -h /tmp | tail -1; nproc")
