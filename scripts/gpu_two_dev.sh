#!/bin/bash
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -x -q -k "two_devices or native_json or facade" 2>&1 | tail -5
