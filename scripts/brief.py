import json, sys
for l in open(sys.argv[1]):
    if l.startswith("{"):
        d = json.loads(l)
        pk = d["roofline"]["per_kernel"]
        print(sys.argv[2], "ms/step", round(d["ms_per_step"], 3), {k: round(v["avg_ms"], 3) for k, v in pk.items()}, "logZ", d["config"]["log_evidence"])
        break
else:
    print(sys.argv[2], "NO RESULT"); print(open(sys.argv[1]).read()[-1500:])
