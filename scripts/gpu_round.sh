#!/bin/bash
# One GPU-box session of the round: tests, bench lines, ncu evidence.  Everything lands under gpurun_out/.
#   TAG=r02a STEPS="tests bench sweep ncu" bash scripts/gpu_round.sh
TAG=${TAG:-r02}
STEPS=${STEPS:-"tests bench sweep ncu"}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi_$TAG.txt 2>&1
for s in $STEPS; do
case $s in
tests)
  timeout 1500 python -m pytest tests -m gpu ${PYTEST_X:--x} -q ${PYTEST_ARGS:-} > $O/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_$TAG.log
  tail -n 15 $O/pytest_$TAG.log ;;
newtests)
  timeout 900 python -m pytest tests/test_fm_fused_gpu.py -m gpu -x -q > $O/pytest_new_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_new_$TAG.log
  tail -n 15 $O/pytest_new_$TAG.log ;;
bench)
  timeout 900 python bench.py --check > $O/bench_${TAG}_fm_c2.json 2> $O/bench_${TAG}_fm_c2.err; echo "bench rc=$?"
  tail -c 600 $O/bench_${TAG}_fm_c2.err ;;
benchquick)
  timeout 600 python bench.py --no-c5 --no-cpu-baseline --steps 100 > $O/bench_${TAG}_fm_c2_quick.json 2> $O/bench_${TAG}_fm_c2_quick.err; echo "bench rc=$?"
  tail -c 600 $O/bench_${TAG}_fm_c2_quick.err ;;
ref)
  timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_${TAG}_reference.json 2> $O/bench_${TAG}_reference.err ;;
sweep)
  for b in 65536 262144; do
    timeout 600 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu-baseline --no-c5 > $O/bench_${TAG}_fm_c2_b$b.json 2> $O/bench_${TAG}_fm_c2_b$b.err
  done ;;
others)
  for w in ffm_c3 nfm_c4 ffm_c5; do
    timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_$w.json 2> $O/bench_${TAG}_$w.err
  done
  for w in ffm_c3 ffm_c5; do
    LCTR_FFM_WARP=0 timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_${w}_cta_per_sample.json 2> $O/bench_${TAG}_${w}_cta_per_sample.err
  done
  LCTR_MLP_UMMA=0 timeout 900 python bench.py --workload nfm_c4 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_nfm_c4_mmasync.json 2> $O/bench_${TAG}_nfm_c4_mmasync.err ;;
ncu)
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 200 --csv --log-file $O/launches_${TAG}_fm_c2.csv \
      python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-c5 > $O/ncu_launch_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fm_fused_kernel|apply_compact" -s 6 -c 40 \
      -o $O/prof_${TAG}_fm_c2 -f python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-c5 > $O/ncu_full_$TAG.log 2>&1
  echo "ncu rc=$?" ;;
dist)
  N=${NGPU:-2}
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 10 \
      > $O/bench_${TAG}_fm_c2_n$N.json 2> $O/bench_${TAG}_fm_c2_n$N.err; echo "dist bench rc=$?"
  tail -c 1500 $O/bench_${TAG}_fm_c2_n$N.err ;;
ncuffm)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ffm_tma_kernel|ffm_fused_kernel" -s 8 -c 2 \
      -o $O/prof_${TAG}_ffm_c3_tma -f python bench.py --workload ffm_c3 --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_ffm_tma_$TAG.log 2>&1
  LCTR_FFM_TMA=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ffm_tma_kernel|ffm_fused_kernel" -s 8 -c 2 \
      -o $O/prof_${TAG}_ffm_c3_old -f python bench.py --workload ffm_c3 --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_ffm_old_$TAG.log 2>&1
  echo "ncu rc=$?" ;;
distq)
  N=${NGPU:-2}
  for v in gpu; do
  LCTR_DIST_FENCE=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 10 --no-c5 \
      > $O/bench_${TAG}_fm_c2_n${N}_$v.json 2> $O/bench_${TAG}_fm_c2_n${N}_$v.err; echo "dist bench $v rc=$?"
  done ;;
umma)
  timeout 120 scripts/lab/umma_lab > $O/umma_lab_$TAG.txt 2>&1; echo "rc=$?" >> $O/umma_lab_$TAG.txt; cat $O/umma_lab_$TAG.txt
  timeout 600 python -m pytest tests/test_mlp_bf16_gpu.py -m gpu -q ${PYTEST_X:--x} > $O/pytest_umma_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_umma_$TAG.log
  tail -n 30 $O/pytest_umma_$TAG.log
  timeout 600 python bench.py --workload nfm_c4 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_nfm_c4.json 2> $O/bench_${TAG}_nfm_c4.err; echo "bench rc=$?"
  LCTR_MLP_UMMA=0 timeout 600 python bench.py --workload nfm_c4 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_nfm_c4_mmasync.json 2> $O/bench_${TAG}_nfm_c4_mmasync.err
  python -c "
import json
for f in ('$O/bench_${TAG}_nfm_c4.json','$O/bench_${TAG}_nfm_c4_mmasync.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('kernels_ms'))
    except Exception as e: print(f, 'ERR', e)
" ;;
ummaq)
  timeout 600 python -m pytest tests/test_mlp_bf16_gpu.py -m gpu -q ${PYTEST_X:--x} > $O/pytest_umma_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_umma_$TAG.log
  tail -n 12 $O/pytest_umma_$TAG.log
  LCTR_MLP_UMMA_TRACE=1 timeout 300 python bench.py --workload nfm_c4 --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2> $O/umma_trace_$TAG.txt; tail -n 3 $O/umma_trace_$TAG.txt
  timeout 600 python bench.py --workload nfm_c4 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_${TAG}_nfm_c4.json 2> $O/bench_${TAG}_nfm_c4.err; echo "bench rc=$?"
  python -c "
import json
for f in ('$O/bench_${TAG}_nfm_c4.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['ms_per_step'], d.get('kernels_ms'))
    except Exception as e: print(f, 'ERR', e)
" ;;
ummaprof)
  LCTR_MLP_UMMA_TRACE=1 timeout 300 python bench.py --workload nfm_c4 --steps 6 --warmup 3 --no-cpu-baseline > /dev/null 2> $O/umma_trace_$TAG.txt; tail -n 8 $O/umma_trace_$TAG.txt
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file $O/launches_${TAG}_nfm_c4.csv \
      python bench.py --workload nfm_c4 --steps 8 --warmup 3 --no-cpu-baseline > $O/ncu_launch_nfm_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nfm_mlp_umma|adagrad_dense" -s 6 -c 4 \
      -o $O/prof_${TAG}_nfm_c4 -f python bench.py --workload nfm_c4 --steps 4 --warmup 3 --no-cpu-baseline > $O/ncu_full_nfm_$TAG.log 2>&1
  echo "ncu rc=$?" ;;
lab)
  bash scripts/lab/run_lab.sh > /dev/null 2>&1; cp $O/lab_b4096.txt $O/lab_${TAG}_b4096.txt; cp $O/lab_b65536.txt $O/lab_${TAG}_b65536.txt ;;
esac
done
ls -la $O | tail -n 30
