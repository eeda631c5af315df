#!/bin/bash
# Marginal SM time of the kernel's stages at full occupancy (development aid): the kernel stops after assembly (1),
# after the inversion (2) or runs to the end (0); the differences are the stages' throughput costs.
for s in 1 2 0; do echo "--- MPC_DEBUG_STOP=$s"; MPC_DEBUG_STOP=$s python tools/occ_sweep.py ${1:-config2} ${2:-65536} | sed -n '1p;2p;4p'; done
