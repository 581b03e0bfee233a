#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_hash_nested.py tests/test_gpu_hash.py -x -q 2>&1 | tail -25 ) > $O/r17_nested.log
tail -25 $O/r17_nested.log
