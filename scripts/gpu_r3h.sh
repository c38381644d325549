#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_json_gpu.py tests/test_sql_aggregate_gpu.py tests/test_inputs_gpu.py -m gpu -x -q 2>&1 | tail -8
