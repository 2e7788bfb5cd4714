import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print("value %.4g  ms/step %.2f us  kernel %.2f us  frac %.3f  e2e %.4g" % (d["value"], d["ms_per_step"]*1e3, d["roofline"]["kernel_ms"]*1e3, d["roofline"]["frac"], d["e2e"]["value"]))
