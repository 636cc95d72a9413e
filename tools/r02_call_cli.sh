#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_lexicon.py tests/test_gpu_host_demo.py -x -q -m gpu 2>&1 | tail -8
