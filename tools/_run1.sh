cd $GRAFT_REPO_ROOT
rm -f gpurun_out/sched4.jsonl
for ch in 56 64; do KV_TOWER_CHUNK=$ch KV_SCHED_CONFIGS=1:0 timeout 200 python tools/bench_sched.py 2>> gpurun_out/sched4.err | sed "s/^/chunk=$ch /" >> gpurun_out/sched4.jsonl; done
cat gpurun_out/sched4.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tower_umma2 -s 3 -c 1 -f -o gpurun_out/prof_tower python tools/bench_net.py --iters 2 > gpurun_out/ncu_tower_full.log 2>&1; echo "ncu full rc=$?"
KV_BENCH_SIMS=8 KV_BENCH_POLICY_PLIES=2 KV_BENCH_COMPARE=0 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_mcts.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 700 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_mcts.json 2> gpurun_out/bench_mcts.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_mcts.err; cat gpurun_out/bench_mcts.json
