#!/bin/bash
# ncu evidence for round 2 (run under gpurun; outputs in gpurun_out/).  Launch lists are cold-cache and serialised:
# compare SHARES.  Full captures: the four projection shapes of encoder layer 1, one packed attention launch of the
# mixed-length batch, LayerNorm / FSMN, and the fp8 projections.
set -x
P="python tools/profile_step.py --batch 32 --steps 1 --warmup 1 --cuda-profiler"
$P > gpurun_out/r02_plain.log 2>&1 || exit 1
$P --mixed > gpurun_out/r02_plain_mixed.log 2>&1 || exit 1
NCU="ncu --clock-control none --profile-from-start off"
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches.csv $P > gpurun_out/r02_ncu1.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches_mixed.csv $P --mixed > gpurun_out/r02_ncu2.log 2>&1
$NCU --set full --import-source on -k regex:k_gemm_tc2 -s 2 -c 4 -f -o gpurun_out/r02_gemm $P > gpurun_out/r02_ncu3.log 2>&1
$NCU --set full --import-source on -k regex:k_attention_tc -s 1 -c 1 -f -o gpurun_out/r02_attn_mixed $P --mixed > gpurun_out/r02_ncu4.log 2>&1
$NCU --set full --import-source on -k regex:k_attention_tc -s 1 -c 1 -f -o gpurun_out/r02_attn $P > gpurun_out/r02_ncu5.log 2>&1
$NCU --set full --import-source on -k "regex:k_layernorm|k_fsmn" -s 3 -c 3 -f -o gpurun_out/r02_ln_fsmn $P > gpurun_out/r02_ncu6.log 2>&1
$NCU --set full --import-source on -k regex:k_gemm_tc2 -s 2 -c 4 -f -o gpurun_out/r02_gemm_fp8 $P --precision fp8 > gpurun_out/r02_ncu7.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -2 gpurun_out/r02_plain.log gpurun_out/r02_plain_mixed.log
