#!/bin/bash
bash scripts/gpu_r2_j.sh
bash scripts/gpu_prof_r02.sh > gpurun_out/prof_r02.log 2>&1; tail -3 gpurun_out/prof_r02.log
