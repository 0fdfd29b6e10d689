#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu --timeout=300 2>&1 | tail -12
