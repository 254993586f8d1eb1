#!/bin/bash
mkdir -p gpurun_out
bash scripts/runs/r2_gpu7.sh
