# A/B of the two-iterations-per-launch kernel inside bench.py (same box, alternating)
i=0
for cfg in "TVL1_T2_WARP0=0" "TVL1_T2_WARP0=1" "TVL1_T2_WARP0=0" "TVL1_T2_WARP0=1"; do
i=$((i+1))
env $cfg timeout 300 python bench.py --quick --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r3o_$i.json 2> gpurun_out/r3o_$i.err
CFG="$cfg" F=gpurun_out/r3o_$i.json python - <<'P'
import json,os
d=json.loads(open(os.environ["F"]).read().strip().splitlines()[-1])
print("%-40s value %.1f ms %.2f e2e %.1f"%(os.environ["CFG"],d["value"],d["ms_per_step"],d["e2e"]["value"]), [(l["level"],round(l["ms"]/d["steps"],1),l["launches"]) for l in d["roofline"]["per_level"][:2]], flush=True)
P
done
