#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/ab_solver.py config2:4096 config2:65536 four_stance:4096 config5:65536 2>&1 | grep -v "inverse \[\|riccati \[" | tee gpurun_out/ab_solver_s12.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_s12.log
