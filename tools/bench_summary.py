import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms_per_step", round(d["ms_per_step"], 2), "quartets/s %.3e" % d["value"], "e2e_ms", round(d["e2e"]["ms_per_step"], 2),
      "diff", d["e2e"]["max_abs_diff_vs_device_path"], "whole", d["roofline"]["whole_build"])
tot = 0
for k, v in d["classes"].items():
    print("%-8s %8.3f ms  screen %6.3f  %6.2f TF/s  q=%d" % (k, v["ms"], v["screen_ms"], v["tflops"], v["quartets"]))
    tot += v["ms"] + v["screen_ms"]
print("sum of class ms", round(tot, 2))
