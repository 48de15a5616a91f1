# usage: bash tools/ab_env_bench.sh VAR val1 val2 ...   (same-box A/B of one environment knob, two rounds)
VAR=$1; shift
for rep in 1 2; do for v in "$@"; do env $VAR=$v python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/abenv_$v.json 2> gpurun_out/abenv_$v.err; python - <<PY
import json
d=json.loads(open("gpurun_out/abenv_$v.json").read().strip().splitlines()[-1])
b=d["roofline"]["step_breakdown_ms"]
print("$VAR=$v", round(d["value"]), round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), {k:b[k] for k in list(b)[:6]})
PY
done; done
