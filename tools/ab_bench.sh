for rep in 1 2; do for v in "$@"; do cp tools/_ab/$v.so fun_asr_gguf_b200/libfunasr_b200.so; python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; python - <<PY
import json
d=json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
b=d["roofline"]["step_breakdown_ms"]
print("$v", round(d["value"]), round(d["ms_per_step"],2), {k:b[k] for k in list(b)[:7]})
PY
done; done
