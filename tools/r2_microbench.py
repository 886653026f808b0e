"""Round-2 probe: shared-memory atomic rates under the access patterns of the intersection kernels."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skbio_b200 import _capi
names = {0: "fp64 DFMA inst/s", 2: "lds64 gathers/s", 3: "atoms random cells 16 warps/SM",
         4: "atoms conflict-free 32 warps/SM", 5: "atoms conflict-free 64 warps/SM",
         6: "atoms random 32 warps/SM", 7: "atoms random 64-bit pair 32 warps/SM",
         8: "atoms conflict-free 64-bit pair 32 warps/SM", 9: "atoms random 64 warps/SM",
         10: "atoms random 64-bit pair 64 warps/SM", 11: "atoms conflict-free 64-bit pair 64 warps/SM"}
out = {}
for k, nm in names.items():
    v = _capi.microbench(k, 0)
    out[nm] = v
    print(f"kind {k:2d} {nm:50s} {v/1e12:8.3f} T/s  = {v/148/1.965e9:6.2f} per clk per SM")
json.dump(out, open("gpurun_out/r2_microbench.json", "w"), indent=1)
