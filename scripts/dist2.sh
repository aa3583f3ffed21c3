cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -k "rank or dist or shim" > gpurun_out/pytest_r02_final_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r02_final_n2.log; tail -5 gpurun_out/pytest_r02_final_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_r02_final_fm_c2_n2.json 2> gpurun_out/bench_r02_final_fm_c2_n2.err; echo "bench rc=$?"
tail -c 400 gpurun_out/bench_r02_final_fm_c2_n2.err
