#!/bin/bash
# Bring-up driver: every group in its own process under a timeout; logs to gpurun_out/.
mkdir -p gpurun_out
LOG=gpurun_out/kernel_check.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
for g in "$@"; do
  echo "##### $g" >> $LOG
  timeout 240 python tools/kernel_check.py --group $g >> $LOG 2>&1
  echo "##### $g exit=$?" >> $LOG
done
grep -E "BAD|exit=|PERF|Error|error" $LOG | tail -120
