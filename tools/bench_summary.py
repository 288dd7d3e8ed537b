import json,sys
for f in sys.argv[1:]:
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, round(d["value"]), round(d["ms_per_step"],3), round(d["e2e"]["value"]), round(d["roofline"]["frac"],3))
    print("   ", " ".join(f'{k}={v["ms_per_step"]:.3f}' for k,v in d["kernels"].items()))
