python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_gpu_mlp.py -x -q -m gpu 2>&1 | tail -4
python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/b_l.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/b_l.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
for s in d['step_roofline']['kernel_ms_per_stage']: print({k:round(v,3) for k,v in s.items()}, round(sum(s.values()),2))
PY
python bench.py --config 2 2>&1 | tail -1 | python -c "import json,sys; d=json.load(sys.stdin); print(d['value'], d['fwd_ms'], d['bwd_ms'])"
