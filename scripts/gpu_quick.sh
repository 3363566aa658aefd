#!/bin/bash
# Quick development check: the pytest selection given as arguments.
timeout 900 python -m pytest "$@" -q 2>&1 | tail -25 > gpurun_out/r2_quick.log; tail -6 gpurun_out/r2_quick.log
