#!/bin/bash
# round 2, call 25 (last GPU minutes): the modules whose aligners share a GPU (arena step-down), at HEAD
cd $GRAFT_REPO_ROOT
timeout 330 python -m pytest tests/test_gpu_parity.py tests/test_python_api.py tests/test_reference_pytests.py -m gpu -x -q > gpurun_out/r2_25_pytest.log 2>&1; tail -4 gpurun_out/r2_25_pytest.log
