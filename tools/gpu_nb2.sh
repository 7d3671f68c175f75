#!/bin/bash
mkdir -p gpurun_out
for f in 1 0; do T2H_FUSE_NB=$f timeout 600 python tools/profile_train_step.py --events --batch 8 2>/dev/null | sed -n '1,4p;/by shape/,$p' > gpurun_out/r2_train_events_nb$f.txt; done
paste -d'|' <(cut -c1-75 gpurun_out/r2_train_events_nb1.txt) <(cut -c1-75 gpurun_out/r2_train_events_nb0.txt) | head -40
