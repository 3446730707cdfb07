#!/bin/bash
# usage: [GPUS=N] scripts/gpurun_retry.sh <outfile> <timeout> <command...>   — retries while the pod answers busy (exit 3 / transient)
out=$1; shift; to=$1; shift
extra=""
if [ -n "$GPUS" ]; then extra="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $extra --timeout $to -- "$@" > $out 2>&1
  rc=$?
  if grep -q "status=transient" $out || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "rc=$rc" >> $out
