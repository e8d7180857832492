#!/bin/bash
# whole TransformerFusion call in train mode on the GPU (train_seq on CudaOps) against the reference's gradients
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_fusion.py tests/test_gpu_train.py tests/test_gpu_trainer.py -q -x -s 2>&1 | tail -n 40
