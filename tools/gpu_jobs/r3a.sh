#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fsi.py -m gpu -x -q -s -k "split" 2>&1 | tail -25
