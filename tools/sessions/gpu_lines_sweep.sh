#!/usr/bin/env bash
# parameter sweep of the lines kernel on the prose text (tools/gpu_prose_exp.py): chunks per turn x idle lanes per hand-out
for turn in 2 3 4 6; do for idle in 1 4 8 16; do
  echo "turn=$turn min_idle=$idle: $(PIRE_B200_LINES_TURN=$turn PIRE_B200_LINES_MIN_IDLE=$idle PROSE_QUICK=1 python tools/gpu_prose_exp.py 2>&1 | grep -E '^(headline|glue10)' | cut -c1-60 | tr '\n' ' ')"
done; done
