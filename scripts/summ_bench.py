import json,sys,glob
for f in sorted(glob.glob(sys.argv[1])):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d["roofline"]["kernels"]
        print(f.split('/')[-1], round(d["value"],1), round(d["ms_per_step"],4), d["config"]["pcg_iterations"], "%.2e"%d["config"]["relres"], round(d["e2e"]["value"],1), d["config"]["mg_levels"], {n:round(v["ms"]*1e3,1) for n,v in k.items()})
    except Exception as e:
        print(f, "ERR", e, open(f).read()[-400:])
