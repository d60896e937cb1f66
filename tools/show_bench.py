"""Print the headline fields of bench.py JSON lines: python tools/show_bench.py gpurun_out/a.json ..."""
import json
import sys

for f in sys.argv[1:]:
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:                                   # noqa: BLE001
        print(f, "ERR", e)
        continue
    s = "%-40s N=%d value %.4g  ms/step %.3f" % (f.split("/")[-1], r.get("n_gpus", 0), r["value"], r["ms_per_step"])
    if "e2e" in r:
        s += "  e2e %.4g" % r["e2e"]["value"]
    if "roofline" in r:
        s += "  kernel %.3f ms frac %.3f" % (r["roofline"].get("kernel_ms", 0), r["roofline"]["frac"])
    if "collector" in r:
        s += "  round %.3f ms (%s)" % (r["collector"]["round_ms"], r["collector"]["mode"])
    if "strong" in r:
        s += "  strong %.4g (%.3f ms)" % (r["strong"]["value"], r["strong"]["ms_per_step"])
    print(s)
