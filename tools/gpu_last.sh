#!/bin/bash
timeout 270 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_last.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_last.log
tail -4 gpurun_out/r02_pytest_last.log
