mkdir -p gpurun_out
timeout 1200 python tools/r2_isect_probe.py u u4 > gpurun_out/r2_isect_probe_f.log 2>&1; cat gpurun_out/r2_isect_probe_f.log | cut -c1-220
