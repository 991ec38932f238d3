#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_api.py -x -q -k graph 2>&1 | grep -E "^E" | cut -c1-900 | head -12
