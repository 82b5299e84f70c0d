#!/bin/bash
# cfg = MULTI:W:CTA:MINB
for cfg in "0:::" "4:::" "4:::2" "4:::3" "4:8:128:4" "4:8:128:3" "4:16:256:1" "2:::2" "2:16:256:3" "4:32:512:1" "4:8:128:2"; do
  IFS=: read M W CTA MINB <<< "$cfg"
  if [ -n "$MINB" ]; then export APTB200_LADDER_MINB=$MINB; else unset APTB200_LADDER_MINB; fi
  echo -n "MULTI=$M "
  APTB200_MULTI=$M C5_W=$W C5_CTA=$CTA timeout 200 python tools/tune_c5.py 2>&1 | tail -1
done
# second sweep
for cfg in "2:8:128:4" "2:8:128:5" "2:8:128:6" "2:16:128:4" "2:8:256:3" "2:16:256:4" "4:8:128:5" "4:16:128:4" "2:16:256:2" "2:32:256:3" "2:4:128:5" "4:4:64:8"; do
  IFS=: read M W CTA MINB <<< "$cfg"
  if [ -n "$MINB" ]; then export APTB200_LADDER_MINB=$MINB; else unset APTB200_LADDER_MINB; fi
  echo -n "MULTI=$M "
  APTB200_MULTI=$M C5_W=$W C5_CTA=$CTA timeout 200 python tools/tune_c5.py 2>&1 | tail -1
done
