#!/bin/bash
timeout 120 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "lbvh" 2>&1 | grep -E "passed|failed|configure with|Error|assert"
