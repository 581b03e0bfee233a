#!/bin/bash
O=gpurun_out; mkdir -p $O
( timeout 1500 python -m pytest tests/test_gpu_more.py -q 2>&1 | tail -40 ) > $O/r11_more.log
( timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_more.py 2>&1 | tail -6 ) > $O/r11_all.log
tail -40 $O/r11_more.log; tail -4 $O/r11_all.log
