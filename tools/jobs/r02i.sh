#!/bin/bash
timeout 400 python tools/graph_debug.py 2>&1 | tail -12
