cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t7.log
tail -4 gpurun_out/t7.log
KV_BENCH_POLICY_PLIES=0 KV_BENCH_INFLIGHT=0 KV_BENCH_COMPARE=0 timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/b7.json 2> gpurun_out/b7.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/b7.json'))
print({k:d[k] for k in ('value','ms_per_step','evals_per_sim','kernels_ms_per_step','clocks')})
print(d['no_cache'], d['random_positions']['value'], d['roofline']['achieved'], d['roofline']['traffic'], d['e2e']['value'])
PY
