#!/bin/bash
# GPU batch F (1 GPU): per-iteration timeline of the default two-q-tile kernel (and its LEAN build) with the tiles in and out of phase
mkdir -p gpurun_out
for ph in 0 900; do
  PF_TL_PHASE=$ph PF_TL_VARIANTS="0x30 0x0c" PF_CHECK_TIMEOUT=100 timeout 150 python tools/gpu_check.py attn4_timeline 2>&1 | grep "attn4_timeline" | grep -v " j=1[89] \| j=2[01] " | sed "s/^/[phase $ph] /"
done
