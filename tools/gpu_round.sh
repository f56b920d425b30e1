#!/bin/bash
# one GPU-box visit: tests, diagnostics, sanitizers, bench (each step bounded by its own timeout)
mkdir -p gpurun_out
T=${1:-r2a}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${T}_pytest.log
tail -5 gpurun_out/${T}_pytest.log
