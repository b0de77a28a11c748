import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f))
        k=d["kernel_ms_per_step"]
        print(f.split('/')[-1], "Msamp/s %.1f Mrays/s %.0f frac %.3f ms/step %.2f | ext %.2f shade %.2f conn %.2f gen %.2f | e2e %.1f" % (d["value"], d["mrays_per_s"], d["roofline"]["frac"], d["ms_per_step"], k["extend"],k["shade"],k["connect"],k["generate"], d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
