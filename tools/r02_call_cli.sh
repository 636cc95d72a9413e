#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_store.py tests/test_gpu_lexicon.py -x -q -m gpu 2>&1 | tail -8
