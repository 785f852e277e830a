import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l)
    if r["op"].startswith("warp"):
        print(r["op"], r["dtype"][:5], "C%d H%d s%.1f"%(r["C"], r["H"], r["sigma"]), "v%d"%r["variant"], "%.3f ms %.0f GB/s frac %.2f" % (r["ms"], r["gbps"], r["frac"]), "x%.1f"%r["speedup"] if r["speedup"] else "")
    else:
        print(r["op"], r["dtype"][:5], "Cd%d Cs%d h%d"%(r["Cd"], r["Cs"], r["h"]), "v%d"%r["variant"], "%.3f ms %.0f GB/s frac %.2f x%.1f" % (r["ms"], r["gbps"], r["frac"], r["speedup"]))
