#!/bin/bash
# gpurun with retry while the pod has no free slot (exit code 3); usage: tools/gr.sh <timeout> '<command>'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" ${GR_GPUS:+--gpus $GR_GPUS} -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
