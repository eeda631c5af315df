#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/ab_solver.py 2>&1 | tee gpurun_out/ab_solver.log
timeout 900 python -m pytest tests -m gpu -x -q -k "forces_match or golden or extreme or status_codes or overflow" 2>&1 | tail -15 | tee gpurun_out/pytest_s4.log
