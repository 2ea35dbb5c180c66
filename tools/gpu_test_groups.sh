#!/bin/bash
# Runs the GPU test groups of the training path in separate processes (a CUDA fault in one group does not poison the
# next) and collects their outputs under gpurun_out/. Usage: gpu_test_groups.sh [group ...] (default: all)
mkdir -p gpurun_out
declare -A G
G[probe]="tests/test_probe_gpu.py"
G[conv]="tests/test_train_kernels_gpu.py -k conv_fwd"
G[bnact]="tests/test_train_kernels_gpu.py -k bn_act"
G[wgrad]="tests/test_train_kernels_gpu.py -k wgrad"
G[prep]="tests/test_train_kernels_gpu.py -k prep_train"
G[chanbwd]="tests/test_train_kernels_gpu.py -k channel_rectifier"
G[fspace]="tests/test_train_kernels_gpu.py -k feat_space"
G[losses]="tests/test_train_kernels_gpu.py -k selfsim"
G[triplet]="tests/test_train_kernels_gpu.py -k triplet"
G[head]="tests/test_train_kernels_gpu.py -k grouped_head"
G[adam]="tests/test_train_gpu.py -k adam"
G[fwd]="tests/test_train_gpu.py -k forward_matches"
G[grads]="tests/test_train_gpu.py -k gradients_match"
G[literal]="tests/test_train_gpu.py -k literal"
G[graph]="tests/test_train_gpu.py -k cuda_graph"
G[step]="tests/test_train_gpu.py -k full_step"
G[ckpt]="tests/test_train_gpu.py -k checkpoint"
G[fullsize]="tests/test_fullsize_gpu.py"
G[recnet]="tests/test_recnet_gpu.py"
G[lfw]="tests/test_lfw_gpu.py"
ORDER="probe conv bnact wgrad prep chanbwd fspace losses triplet head adam fwd grads literal graph step ckpt fullsize lfw"
[ $# -gt 0 ] && ORDER="$*"
for name in $ORDER; do
  timeout 900 python -m pytest ${G[$name]} -m gpu -q -s > gpurun_out/t_$name.log 2>&1
  echo "== $name rc=$? $(tail -1 gpurun_out/t_$name.log)"
done
