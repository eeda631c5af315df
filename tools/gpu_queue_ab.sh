set -u
cd /root/repo
echo "== dynamic queue"; timeout 600 python tools/ab_solver.py config2:4096 config2:65536 config5:65536 four_stance:4096 2>&1 | grep -v "inverse \[\|riccati \["
echo "== static stride"; MPC_RIC_STATIC=1 timeout 600 python tools/ab_solver.py config2:4096 config2:65536 config5:65536 2>&1 | grep -v "inverse \[\|riccati \["
python bench.py --config 2 --steps 100 --warmup 5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/q_c2.json; python tools/show_bench.py gpurun_out/q_c2.json | grep -v parity
timeout 900 python -m pytest tests -m gpu -x -q -k "full_size_properties or mixed_stance or overflow or two_device_slots" 2>&1 | tail -3
