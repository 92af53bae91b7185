timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -4
for i in 1 2; do
NPP_WG_NO_BALANCE=1 timeout 120 python tests/diag_step_time.py 2>&1 | tail -3
timeout 120 python tests/diag_step_time.py 2>&1 | tail -3
done
