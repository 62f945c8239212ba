#!/bin/bash
# ncu captures of one round at the BENCH shape (4-sample north_star step); run under gpurun on ONE GPU:
#   gpurun --timeout 1200 -- 'bash tools/ncu_round.sh r02'
# Summaries are made afterwards on the CPU box with tools/ncu_summary.py and committed under profiles/.
set -x
TAG=${1:-r02}
OUT=gpurun_out
NCU="ncu --clock-control none"
# every launch of 1 warm-up + 1 measured step (plan building included; shares only -- cold cache, serialised)
$NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $OUT/launches_$TAG.csv python tools/profile_step.py 4 2 > $OUT/ncu_launches_$TAG.log 2>&1
# per forward: 99 conv_gemm launches (0 stem, 1-6 layer1, ... 29-33 FAX stage-0 token GEMMs, 34 out-projection),
# 12 token_mlp launches (0, 1 = stage 0), 13 window_attn_v3 launches (0 = stage-0 attn1)
FULL="$NCU --set full --import-source on"
$FULL -k regex:window_attn_v3 -s 39 -c 2 -o $OUT/prof_attn_$TAG -f python tools/profile_step.py 4 4 > $OUT/ncu_attn_$TAG.log 2>&1
$FULL -k regex:conv_gemm_kernel -s 298 -c 2 -o $OUT/prof_conv3x3_$TAG -f python tools/profile_step.py 4 4 > $OUT/ncu_conv_$TAG.log 2>&1
$FULL -k regex:conv_gemm_kernel -s 326 -c 3 -o $OUT/prof_token_$TAG -f python tools/profile_step.py 4 4 > $OUT/ncu_token_$TAG.log 2>&1
$FULL -k regex:token_mlp -s 36 -c 1 -o $OUT/prof_mlp_$TAG -f python tools/profile_step.py 4 4 > $OUT/ncu_mlp_$TAG.log 2>&1
ls -la $OUT/*.ncu-rep | tail -6
