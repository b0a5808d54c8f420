set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_mcts.py tests/test_gpu_net.py -m gpu -x -q > gpurun_out/t2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t2.log
tail -5 gpurun_out/t2.log
KV_BENCH_POLICY_PLIES=0 KV_BENCH_INFLIGHT=0 timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/b2.json 2> gpurun_out/b2.err; echo "bench rc=$?"
tail -3 gpurun_out/b2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/b2.json'))
print({k:d[k] for k in ('value','ms_per_step','evals_per_sim','kernels_ms_per_step','clocks')})
print(d['schedule']['single_stream'], d['no_cache'], d['random_positions'], d['roofline']['achieved'], d['e2e']['value'])
PY
