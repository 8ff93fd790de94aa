# A/B inside bench.py (same box, alternating): arguments = environment settings to compare
i=0
for cfg in "$@"; do
i=$((i+1))
env $cfg timeout 300 python bench.py --quick --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
CFG="$cfg" F=gpurun_out/ab_$i.json python - <<'P'
import json,os
d=json.loads(open(os.environ["F"]).read().strip().splitlines()[-1])
fb=[(l["level"],round(l["first_block"]["ms"]/d["steps"],2)) for l in d["roofline"]["per_level"] if l.get("first_block")]
print("%-40s value %.1f ms %.2f e2e %.1f"%(os.environ["CFG"],d["value"],d["ms_per_step"],d["e2e"]["value"]), [(l["level"],round(l["ms"]/d["steps"],1),l["launches"]) for l in d["roofline"]["per_level"][:2]], "first blocks", fb, "t1 frac %.3f"%d["roofline"]["frac"], flush=True)
P
done
