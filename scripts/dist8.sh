# FM C2 weak scaling + FFM C5 global-batch split at N GPUs (bench.py under torchrun); outputs under gpurun_out/
N=${NGPU:-8}
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 50 --warmup 5 \
    > gpurun_out/bench_r02_fm_c2_n$N.json 2> gpurun_out/bench_r02_fm_c2_n$N.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench_r02_fm_c2_n$N.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r02_fm_c2_n$N.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], (d.get('c5') or {}).get('value'))"
