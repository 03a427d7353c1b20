#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "backward or sigma_noise" -s 2>&1 | grep -E "passed|failed|Error|assert|worst|snb" | head -20 | cut -c1-500
