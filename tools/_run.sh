(timeout 600 python -m pytest tests/test_gpu_pnp.py -m gpu -x -q -s 2>&1 | tail -15) > gpurun_out/s7_pnp.log 2>&1; cat gpurun_out/s7_pnp.log
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/s7_pytest.log 2>&1; tail -3 gpurun_out/s7_pytest.log
