#!/bin/bash
timeout 1400 python -m pytest tests -x -q -m gpu --timeout=300 --durations=8 2>&1 | tail -25
