#!/bin/bash
mkdir -p gpurun_out
bash scripts/r2_gpu7.sh
