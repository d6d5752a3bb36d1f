import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        r=d["roofline"]; print("%s: %s value %.3e rows/s  ms/step %.3f  kernel %.4f ms x%d  %.0f GB/s frac %.3f  e2e %.3e (%.1f GB/s)  launches %d"%(f,d["config"]["workload"],d["value"],d["ms_per_step"],r["launch_ms"],d["config"]["chunks_per_step"],r["achieved"],r["frac"],d["e2e"]["value"],d["e2e"]["h2d_gb_per_s"],d["gpu_launches"]))
    except Exception as e: print(f,"ERR",e)
