timeout 100 python tools/run_adjoint.py rd 2>&1 | grep -v Warn; timeout 100 python tools/run_adjoint.py bf 2>&1 | grep -v Warn
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/b.json; python - <<'PY'
import json
d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print(d["value"], d["roofline"]["frac"], d["e2e"]["value"]); print({k:round(v["ms"],4) for k,v in d["extras"].items() if "ms" in v})
PY
