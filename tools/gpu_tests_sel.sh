#!/bin/bash
# selected GPU tests (arguments = pytest -k expression), printing the tests' own diagnostics
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -k "$1" --timeout 900 > gpurun_out/pytest_sel.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_sel.log
grep -v "^$" gpurun_out/pytest_sel.log | tail -40
