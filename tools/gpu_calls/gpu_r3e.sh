#!/bin/bash
# ncu --set full (with source) of the current STFT kernel at configs[3]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
timeout 240 ncu --set full --clock-control none --import-source on -k regex:stft_mel_kernel -s 2 -c 1 -o $O/r3e_stft -f python tools/prof_kernels.py --only stft_mel_c4 --iters 1 > $O/r3e_ncu.log 2>&1
tail -3 $O/r3e_ncu.log; ls -la $O/r3e*
