import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); print(f, d["value"], d["unit"], d["ms_per_step"], "ms  e2e", round(d["e2e"]["value"],2) if d.get("e2e") else None, "launches", d["gpu_launches"])
        print("  ", {k:(v["ms"],v.get("tflops")) for k,v in d["breakdown"].items() if k!="_step"})
        print("  ", d["breakdown"]["_step"], d["clocks"], d.get("roofline"))
    except Exception as e: print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-1500:])
