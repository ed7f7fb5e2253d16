#!/bin/bash
for mb in 2 0; do echo "META_BUFS=$mb"; B200SP_TC_META_BUFS=$mb timeout 120 python tools/dev_split.py 2>&1 | grep rows; done
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu 2>&1 | tail -3
