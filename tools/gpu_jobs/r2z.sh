#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_time_loop.py -m gpu -x -q -k "monaghan_kajtar_wall or boundary_model_variants" 2>&1 | tail -15
