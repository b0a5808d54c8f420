cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_net.py tests/test_gpu_mcts.py tests/test_gpu_train.py -m gpu -x -q > gpurun_out/t5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t5.log
tail -6 gpurun_out/t5.log
KV_TOWER_FUSED=0 timeout 300 python -m pytest tests/test_gpu_net.py tests/test_gpu_mcts.py -m gpu -x -q > gpurun_out/t5b.log 2>&1; echo "pytest(per-layer) rc=$?" >> gpurun_out/t5b.log
tail -3 gpurun_out/t5b.log
rm -f gpurun_out/sched3.jsonl
for cfg in "74 1:0" "74 0:0" "74 1:1" "148 1:0"; do set -- $cfg; KV_TOWER_CHUNK=$1 KV_SCHED_CONFIGS=$2 timeout 200 python tools/bench_sched.py 2>> gpurun_out/sched3.err | sed "s/^/chunk=$1 /" >> gpurun_out/sched3.jsonl; done
cat gpurun_out/sched3.jsonl; tail -3 gpurun_out/sched3.err
