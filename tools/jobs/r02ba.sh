#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
