#!/bin/bash
# runs the kernel lab on the GPU box: scripts/lab/run_lab.sh  (outputs under gpurun_out/)
mkdir -p gpurun_out
python scripts/lab/dump_batch.py /tmp/lab 1000000 4096 65536 > gpurun_out/lab_dump.txt 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/lab_smi.txt 2>&1
timeout 300 scripts/lab/fm_lab /tmp/lab_F1000000_B4096.bin > gpurun_out/lab_b4096.txt 2>&1
echo "rc=$?" >> gpurun_out/lab_b4096.txt
timeout 300 scripts/lab/fm_lab /tmp/lab_F1000000_B65536.bin > gpurun_out/lab_b65536.txt 2>&1
echo "rc=$?" >> gpurun_out/lab_b65536.txt
tail -n 60 gpurun_out/lab_b4096.txt
